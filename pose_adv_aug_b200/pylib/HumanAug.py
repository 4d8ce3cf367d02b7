"""The part of the reference `pylib/HumanAug.py` that sits on the validation path next to the network
(stack-hg.py:225-232): horizontal-flip test-time augmentation of the heat-maps, on the GPU.

  flip_channels(maps)                                ref pylib/HumanAug.py:199-210
  shuffle_channels_for_horizontal_flipping(maps)     ref pylib/HumanAug.py:179-197
  flip_merge(output1, output2)                       = (output1 + shuffle(flip(output2))) / 2 in ONE kernel

The reference does the flip with numpy on a host copy of the output; here the W-flip, the six left/right channel
swaps and the mean are the index arithmetic of a single pass (`hgk_flip_merge_nchw`).  The image crop / warp
functions of the reference file (disk + PIL/scipy data pipeline) are out of scope.
"""
import ctypes

import torch

from .._lib import get_lib, HGKError

FLIP_PAIRS = [[1, 4], [0, 5], [12, 13], [11, 14], [10, 15], [2, 3]]      # ref:182


def _check(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise HGKError("HumanAug.%s runs on CUDA tensors only (no CPU fallback)" % what)
    if t.dim() not in (3, 4):
        raise ValueError('tensor dimension is not right')                 # ref:188,206
    return t.contiguous().float()


def _launch(a, b, pairs, flip_w):
    """out = a is None ? f(b) : (a + f(b)) / 2,  f = channel swaps after an optional W-flip."""
    lib = get_lib()
    squeeze = b.dim() == 3
    b4 = b.unsqueeze(0) if squeeze else b
    a4 = None if a is None else (a.unsqueeze(0) if squeeze else a)
    if a4 is not None and a4.shape != b4.shape:
        raise ValueError("flip merge: shapes differ (%s vs %s)" % (tuple(a4.shape), tuple(b4.shape)))
    N, C, H, W = b4.shape
    flat = [int(v) for p in pairs for v in p]
    arr = (ctypes.c_int * max(len(flat), 1))(*flat)                      # host array, read during the call only
    out = torch.empty_like(b4)
    lib.check(lib.flip_merge_nchw(0 if a4 is None else a4.data_ptr(), b4.data_ptr(), N, C, H, W,
                                  ctypes.cast(arr, ctypes.c_void_p).value, len(pairs), int(flip_w), out.data_ptr(),
                                  torch.cuda.current_stream(b.device).cuda_stream), "hgk_flip_merge_nchw")
    return out[0] if squeeze else out


def flip_merge(output1, output2, pairs=FLIP_PAIRS):
    """(output1 + shuffle_channels_for_horizontal_flipping(flip_channels(output2))) / 2, stack-hg.py:229-232.
    output1: heat-maps of the images, output2: heat-maps of the W-flipped images; NCHW CUDA tensors."""
    return _launch(_check(output1, "flip_merge"), _check(output2, "flip_merge"), pairs, True)


def flip_channels(maps):
    """ref:199-210 -- horizontally flip every channel (returns a new float tensor)."""
    return _launch(None, _check(maps, "flip_channels"), [], True)


def shuffle_channels_for_horizontal_flipping(maps):
    """ref:179-197 -- swap the left/right joint channels IN PLACE (as the reference does) and return `maps`."""
    out = _launch(None, _check(maps, "shuffle_channels_for_horizontal_flipping"), FLIP_PAIRS, False)
    maps.copy_(out)
    return maps
