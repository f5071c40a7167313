"""Dev tool: where the wall time of PSpinor.imaginary() goes at 2048^2 (cProfile, host side)."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

unwrap = sys.argv[1] if len(sys.argv) > 1 else 'none'
q = bench.build_problem(2048, tag='w')
q.imaginary(1 / 50, 2, 'cuda', unwrap=unwrap)
q = bench.build_problem(2048, tag='p')
pr = cProfile.Profile()
pr.enable()
q.imaginary(1 / 50, 20, 'cuda', unwrap=unwrap)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
