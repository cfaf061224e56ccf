"""One GPU's share of BASELINE config 4 (S=100, E=100 000, 512 chains): score kernel and MCMC iteration time (diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import bench
import nem_b200 as nb
print(json.dumps(bench.config4_shape_leg(nb, nb.synth)))
