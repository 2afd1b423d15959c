#!/bin/bash
for v in "SB_X=1" "SB_LABEL_PLAIN=1"; do echo "== $v"; env $v python scripts/prof_hbm.py 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l)
        if 'label' in r['kernel'] or 'zero' in r['kernel']: print('%-44s %7.1f GB/s  frac %.3f  %.3f ms' % (r['kernel'], r['achieved'], r['frac'], r['ms']))"; done
