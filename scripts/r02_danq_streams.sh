for n in 1 2 3 4; do python scripts/danq_config4.py 1500 $n 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('streams $n', round(d['device_resident']['mutants_per_s']), round(d['device_resident']['ms_per_sequence'],3), round(d['to_host']['mutants_per_s']))"; done
