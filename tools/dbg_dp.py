import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import hg_oracle as O, synth
from pose_adv_aug_b200 import dist as hdist, HourglassTrainer
from pose_adv_aug_b200.models import asn_stacked_hg as M
def log(*a):
    print("[r%s %.1f]" % (os.environ.get("RANK"), time.time() % 1000), *a, flush=True)
log("start")
rank, world, local = hdist.init_from_env()
log("init done", rank, world, local)
dev = torch.device("cuda", local)
t = torch.ones(4, device=dev); dist.all_reduce(t); torch.cuda.synchronize(); log("allreduce ok", t[0].item())
S, C, N, R = 2, 64, 4, 128
net = M.create_hg(S, 1, 16, C)
use_graph = len(sys.argv) > 1 and sys.argv[1] == "graph"
tr = HourglassTrainer(net, N, R, device=dev, use_graph=use_graph)
log("trainer built")
x = synth.make_images(N, R, seed=60 + rank).to(dev); tt = synth.make_heatmaps(N, R, 16, seed=70 + rank).to(dev)
for i in range(3):
    l = float(tr.step(x, tt)); log("step", i, l)
dist.barrier(); log("done")
dist.destroy_process_group()
