import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch as th, numpy as np
import test_gpu_garf as T
cuda = th.device("cuda:0")
g = T._g()
for prec in ("fp32", "tf32", "bf16"):
    prop, rad = T._seeded_nets(cuda)
    prop.matmul_precision = rad.matmul_precision = prec
    rgb, dens = rad(g["net_pos"].to(cuda), g["net_dir"].to(cuda))
    ((rgb * g["up_rgb"].to(cuda)).sum() + (dens * g["up_density"].to(cuda)).sum()).backward()
    errs = []
    for n, p in rad.named_parameters():
        ref = g["rad.grad." + n]
        errs.append(float((T._thin(p.grad).cpu() - ref).norm() / (ref.norm() + 1e-12)))
    print(prec, "rgb max abs", float((rgb.cpu() - g["rad_rgb"]).abs().max()), "dens max rel", float(((dens.cpu() - g["rad_density"]).abs() / (g["rad_density"].abs() + 1e-3)).max()),
          "grad rel-L2 max", max(errs), "median", float(np.median(errs)))
