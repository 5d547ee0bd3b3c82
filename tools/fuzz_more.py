"""more draws of the combination fuzz (tests/test_gpu_fuzz.py) than the suite runs: python
tools/fuzz_more.py [first_seed] [n]  -- prints one line per draw, raises on the first failure"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import test_gpu_fuzz as F
first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
for s in range(first, first + n):
    print('seed', s, flush=True)
    F.test_scenarios_constraints_and_all_parameters_together(s)
    F.test_tracking_with_scenarios_constraints_and_all_parameters(s + 50000, 1750 + (s * 37) % 300)
    F.test_random_biome_configurations_with_constraints(s + 90000)
print("fuzz_more: %d draws of each kind passed" % n)
