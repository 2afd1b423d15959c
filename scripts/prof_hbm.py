"""Launches each HBM-bound kernel of stage 1 / 2a a few times on bench.py's shapes, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:'window_kernel|label_bins|ism_expand' -o gpurun_out/prof_hbm python scripts/prof_hbm.py --ncu
"""
import json, os, sys
sys.path.insert(0, os.getcwd())
import bench as B
import torch
from selene_b200.sequences import Genome
names, seqs = B.synth_genome_arrays(240.0)
G = Genome.from_sequences(names, seqs)
Genome.update_bases_order(['A', 'C', 'G', 'T'])
if '--ncu' in sys.argv:
    B._timed.__defaults__ = (0, 2)      # under ncu: no warm-up, two launches each
for r in B.hbm_rooflines(G, names, seqs, 6549.4):
    print(json.dumps(r))
