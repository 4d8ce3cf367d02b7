import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from pose_adv_aug_b200 import HourglassTrainer, synth
from pose_adv_aug_b200.models import asn_stacked_hg as M
dev = torch.device("cuda", 0)
net0 = M.create_hg(2, 1, 16, 256)
sd = synth.make_state_dict(synth.schema_of(net0), seed=1, perturb_bn=False)
x = synth.make_images(24, 256, seed=2).to(dev); t = synth.make_heatmaps(24, 256, 16, seed=3).to(dev)
for ns, nl in [tuple(int(v) for v in a.split(",")) for a in (sys.argv[1:] or ["8,3", "10,3"])]:
    net = M.create_hg(2, 1, 16, 256); net.load_state_dict(sd)
    tr = HourglassTrainer(net, 24, 256, device=dev, n_streams=ns, n_low=nl)
    tr.x.copy_(x); tr.t.copy_(t)
    for _ in range(5): tr.step_resident()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40): tr.step_resident()
    e1.record(); torch.cuda.synchronize()
    print("streams %2d low %d: %.3f ms/step" % (ns, nl, e0.elapsed_time(e1) / 40), flush=True)
    del tr, net
    torch.cuda.empty_cache()
