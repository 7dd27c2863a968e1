import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch as th
import test_gpu_mlp as T
cuda = th.device("cuda:0")
for case in [(300, 1, 128, True, False, 1, False, False), (300, 1, 64, True, False, 1, False, True), (128*5+17, 4, 256, True, False, 2, True, True)]:
    errs = T._grad_case(cuda, *case, seed=2)
    print(case)
    for k, v in errs.items():
        print(f"   {k:32s} {v:.4f}")
