import cProfile, pstats, sys, io
sys.path.insert(0, "/root/repo")
sys.argv = ["x", "4", "1"]
import runpy
# warm up by running the bench script once, then profile a few steps
g = runpy.run_path("/root/repo/tools/bench_train_step.py")
m, lr_t, hr_t = g["m"], g["lr_t"], g["hr_t"]
import torch
def step():
    m.run_gradient_descent(lr_t, hr_t, m.generator_weights, weight_gen_advers=1e-3, train_gen=True, train_disc=False)
    m.run_gradient_descent(lr_t, hr_t, m.discriminator_weights, weight_gen_advers=1e-3, train_gen=False, train_disc=True)
    torch.cuda.synchronize()
step()
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
