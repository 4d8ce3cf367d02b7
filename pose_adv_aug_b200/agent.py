"""On-GPU sampling of the agent's augmentation choices.

The reference turns the agent's logits into probabilities with softmax, copies them to the host and calls
`np.random.choice(K, 1, p=probs[j])` per sample, scale then rotation (joint-train-pose-s-r-agent.py:252-271 and
:344-363) -- a D2H sync in the middle of config 3's critical path.  Here the uniforms are drawn from numpy's global
RandomState BEFORE the forward result is needed (same stream consumption: one uniform per `choice`, interleaved
scale/rotation per sample), copied H2D asynchronously, and `hgk_softmax_sample` does softmax + inverse-CDF search
on the device.  With the same `np.random.seed` the indices equal the reference's (the uniform falls within one fp32
rounding of a CDF step with probability ~1e-7 per draw).
"""
import numpy as np
import torch

from ._lib import get_lib, HGKError


def draw_uniforms(n, rng=None):
    """[n,2] fp64 uniforms in the order the reference consumes them (sample-major, scale then rotation)."""
    r = np.random if rng is None else rng
    return torch.from_numpy(np.asarray(r.random_sample(2 * n), dtype=np.float64).reshape(n, 2))


def softmax_sample(logits, u):
    """logits [N,K] CUDA fp32, u [N] fp64 uniforms (host or device) -> (probs [N,K] fp32, index [N] int64), on device."""
    if not isinstance(logits, torch.Tensor) or not logits.is_cuda:
        raise HGKError("agent.softmax_sample runs on CUDA tensors only (no CPU fallback)")
    lib = get_lib()
    l = logits.detach().contiguous().float()
    N, K = l.shape
    ud = u.to(device=l.device, dtype=torch.float64, non_blocking=True).contiguous()
    if ud.numel() != N:
        raise ValueError("need one uniform per row")
    probs = torch.empty_like(l)
    index = torch.empty(N, device=l.device, dtype=torch.int64)
    lib.check(lib.softmax_sample(l.data_ptr(), N, K, ud.data_ptr(), probs.data_ptr(), index.data_ptr(),
                                 torch.cuda.current_stream(l.device).cuda_stream), "hgk_softmax_sample")
    return probs, index


def sample_scale_rotation(pred_scale_distri, pred_rotation_distri, uniforms=None, rng=None):
    """The sampling block of joint-train-pose-s-r-agent.py:252-271: returns
    (scale_probs, rotation_probs, scale_index, rotation_index), all CUDA tensors (no host round trip)."""
    n = pred_scale_distri.shape[0]
    u = draw_uniforms(n, rng) if uniforms is None else uniforms
    if u.device.type == "cpu":
        u = u.pin_memory() if torch.cuda.is_available() else u
    ps, si = softmax_sample(pred_scale_distri, u[:, 0])
    pr, ri = softmax_sample(pred_rotation_distri, u[:, 1])
    return ps, pr, si, ri


# ---------------------------------------------------------------------------------------------------------------------
# load_batch_data on the GPU: the agent's sampled (scale, rotation) applied to resident images
# ---------------------------------------------------------------------------------------------------------------------
def sample_from_small_gaussian(mean, var, rng=None):
    """ref data/joint_train_s_r_agent.py:15-16"""
    r = np.random if rng is None else rng
    return max(mean - var + 1e-3, min(mean + var, mean + r.randn() * var))


class AgentBatchLoader(object):
    """`load_batch_data` of joint-train-pose-s-r-agent.py:425-450 without the DataLoader: the index-list path of
    `AGENT.__getitem__` (data/joint_train_s_r_agent.py:100-175, separate_s_r = False) for a whole batch, from images that
    stay resident in HBM.

    images: list of float32 CUDA tensors [3,H,W] in [0,1] (what `imutils.load_image` returns, uploaded once);
    annos:  list of dicts with 'joint_self' [16,3], 'objpos' [2], 'scale_provided', 'normalizer' (the MPII json entries).
    The host does per sample what costs nothing (a dozen scalars: the Gaussian draws, the flip coin, the colour gains, the crop
    window, the 16 transformed joints); the pixels -- flip, colour gain + clamp, byte-scaling, PIL-exact rotate / resize, the
    heat-map blobs -- never leave the GPU.  Random numbers are consumed from numpy's global RandomState in the order
    `__getitem__` consumes them (scale draw, rotation draw, flip coin, three colour gains per sample)."""

    MATCHED = ([0, 5], [1, 4], [2, 3], [10, 15], [11, 14], [12, 13])          # ref pylib/HumanAug.py:240-243

    def __init__(self, images, annos, inp_res=256, out_res=64, std_size=200):
        if len(images) != len(annos):
            raise ValueError("one annotation per image")
        for im in images:
            if not isinstance(im, torch.Tensor) or not im.is_cuda or im.dtype != torch.float32 or im.dim() != 3 or im.shape[0] != 3:
                raise HGKError("AgentBatchLoader keeps float32 CUDA images [3,H,W] (no CPU fallback)")
        self.images, self.annos = images, annos
        # per annotation, once: the float32 arrays __getitem__ builds from the json entry on every call (ref :105-124)
        f32 = np.float32
        self._prep = []
        for a in annos:
            s = f32(a['scale_provided'])
            c = np.asarray(a['objpos'], dtype=np.float32).copy()
            c[1] = c[1] + f32(15) * s                                          # ref :122-124 (torch float32 arithmetic)
            self._prep.append((np.asarray(a['joint_self'], dtype=np.float32)[:, 0:2].copy(), c, s * f32(1.25), a['normalizer'] * 0.6))
        self._perm = np.arange(16)
        for i, j in self.MATCHED:
            self._perm[[i, j]] = self._perm[[j, i]]
        self.inp_res, self.out_res, self.std_size = inp_res, out_res, std_size
        self.scale_means = np.arange(-0.6, 0.61, 0.2)                          # ref :31-35
        self.scale_var = 0.05
        self.rotation_means = np.arange(-60, 61, 20)
        self.rotation_var = 5

    def sample_params(self, a, scale_index, rotation_index, width, rng=None):
        """The host half of __getitem__ for one sample: returns (pts [16,2] f32, c [2] f32, s_aug f32, r_aug, normalizer,
        flip, gains)."""
        r = np.random if rng is None else rng
        f32 = np.float32
        if isinstance(a, (int, np.integer)):                                   # index into the annotations prepared in __init__
            pts, c, s, normalizer = self._prep[a]
            pts, c = pts.copy(), c.copy()
        else:
            pts = np.asarray(a['joint_self'], dtype=np.float32)[:, 0:2].copy()
            c = np.asarray(a['objpos'], dtype=np.float32).copy()
            s = f32(a['scale_provided'])
            c[1] = c[1] + f32(15) * s                                          # ref :122-124 (torch float32 arithmetic)
            s = s * f32(1.25)
            normalizer = a['normalizer'] * 0.6
        scale_factor = sample_from_small_gaussian(self.scale_means[scale_index], self.scale_var, r)       # ref :139-143
        r_aug = sample_from_small_gaussian(self.rotation_means[rotation_index], self.rotation_var, r)
        s_aug = s * f32(2 ** scale_factor)
        flip = bool(r.random() <= 0.5)                                         # ref :160
        if flip:
            pts[:, 0] = f32(width) - pts[:, 0]
            pts = pts[self._perm]                                              # the six left / right swaps of shufflelr
            c[0] = f32(width) - c[0]
        gains = [r.uniform(0.6, 1.4) for _ in range(3)]                        # ref :166-168
        return pts, c, s_aug, r_aug, normalizer, flip, gains

    def load_batch(self, scale_index_list, rotation_index_list, img_index, rng=None):
        """-> (img [N,3,res,res], heatmap [N,16,out,out]) on the GPU, and c [N,2], s [N,1], r [N,1], grnd_pts [N,16,2],
        normalizer [N] as host tensors, like the tuple the reference's DataLoader yields."""
        from .pylib import HumanAug, HumanPts
        idx = [int(i) for i in (img_index.tolist() if isinstance(img_index, torch.Tensor) else img_index)]
        n = len(idx)
        crops, cs, ss, rs, ptss, norms, pts_aug, flips, gainss = [], [], [], [], [], [], [], [], []
        for k, i in enumerate(idx):
            img = self.images[i]
            pts, c, s_aug, r_aug, normalizer, flip, gains = self.sample_params(
                i, int(scale_index_list[k]), int(rotation_index_list[k]), int(img.shape[2]), rng)
            # flip, colour gain + clamp and the CHW -> HWC view are evaluated on load by the batched crop kernels
            crops.append(img); flips.append(flip); gainss.append(gains)
            cs.append(c); ss.append(s_aug); rs.append(r_aug); ptss.append(pts); norms.append(normalizer)
            pa = HumanAug.TransformPts(pts, c, np.array([s_aug], dtype=np.float32), r_aug, self.out_res, self.std_size)
            pa[(pts[:, 0] <= 0) | (pts[:, 1] <= 0)] = 0                        # ref :207-210
            pts_aug.append(pa)
        inp = HumanAug.crop_batch_resident(crops, flips, gainss, np.stack(cs), np.array(ss, dtype=np.float32), np.array(rs),
                                           self.inp_res, self.std_size)
        # the renderer takes float32 points while the reference truncates the float64 ones (draw_gaussian: int(pt -+ 3)):
        # hand it, per coordinate, a float32-exact stand-in with the same validity and the same truncations -- the value
        # itself when it is an integer, the middle of its unit interval otherwise
        pa = np.stack(pts_aug)
        pa = np.where((pa <= 0) | (pa == np.floor(pa)), pa, np.floor(pa) + 0.5)
        pa = torch.from_numpy(pa).to(device=inp.device, dtype=torch.float32)
        heat, _ = HumanPts.pts2heatmap(pa, [self.out_res, self.out_res], sigma=1)
        return (inp, heat, torch.from_numpy(np.stack(cs)), torch.from_numpy(np.array(ss, dtype=np.float32)).view(n, 1),
                torch.tensor(rs, dtype=torch.float32).view(n, 1), torch.from_numpy(np.stack(ptss)), torch.tensor(norms, dtype=torch.float64))
