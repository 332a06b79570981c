"""Host time of one ChainLoss forward+backward (cfg3 shapes): is the step bound by the host issuing launches?
Prints the asynchronous host time per step, the device time per step, and the cProfile top of the host path."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pychain_b200 import ChainLoss, synth
dev = torch.device("cuda:0")
spec = synth.workload_spec("cfg3", variant="uniform")
graph, num_batch, x_host, lengths = synth.build_workload(spec, 0, spec["B"])
x = x_host.to(dev)
crit = ChainLoss(graph, 1e-5, avg=False)


def step():
    xg = x.detach().requires_grad_(True)
    loss = crit(xg, lengths, num_batch)
    loss.backward()
    return loss


for _ in range(10):
    step()
torch.cuda.synchronize()
n = 200
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(n):
    step()
t1 = time.perf_counter(); e1.record()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host issue time per step %.3f ms | device time per step %.3f ms | wall incl. drain %.3f ms" % (
    (t1 - t0) / n * 1e3, e0.elapsed_time(e1) / n, (t2 - t0) / n * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr, stream=sys.stdout)
st.sort_stats("cumulative").print_stats(28)
