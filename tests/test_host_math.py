"""CPU: the algebra the fused kernels rely on, stated with plain torch / Python (no kernel runs here).

* neck: conv1x1(cat[upsample2x(a), b]) == upsample2x(W_a . a) + W_b . b  before bias / activation -- what
  yms_conv_plan_add_upsampled computes (reference: Upsample.forward components.py:159-160, torch.cat in Neck.forward
  yolov8_neck.py:77-83, C2f.conv1 components.py:108);
* NMS: the normalised bitonic network of nms.cu::segment_sort (mirror step + half-cleaners, every compare-exchange puts the
  minimum at the lower index) sorts a segment of ANY length in place when indices >= len are treated as +inf.
"""
import random

import torch
import torch.nn.functional as F


def test_conv1x1_over_upsampled_concat_is_linear_in_its_two_halves():
    g = torch.Generator().manual_seed(0)
    a = torch.randn(2, 48, 5, 7, generator=g, dtype=torch.float64)
    b = torch.randn(2, 24, 10, 14, generator=g, dtype=torch.float64)
    w = torch.randn(32, 72, 1, 1, generator=g, dtype=torch.float64)
    bias = torch.randn(32, generator=g, dtype=torch.float64)
    want = F.silu(F.conv2d(torch.cat([F.interpolate(a, scale_factor=2, mode="nearest"), b], 1), w, bias))
    low = F.conv2d(a, w[:, :48])                                       # half resolution, no bias
    got = F.silu(F.interpolate(low, scale_factor=2, mode="nearest") + F.conv2d(b, w[:, 48:]) + bias.view(1, -1, 1, 1))
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)


def _segment_sort(s):
    """Index logic of nms.cu::segment_sort, one 'thread' at a time."""
    n = len(s)
    if n < 2:
        return
    p = 2
    while p < n:
        p <<= 1
    half_pairs = p >> 1
    k, lg = 2, 0
    while k <= p:
        for t in range(half_pairs):                                    # mirror step
            blk, off = t >> lg, t & ((k >> 1) - 1)
            i, l = blk * k + off, blk * k + (k - 1 - off)
            if l < n and s[i] > s[l]:
                s[i], s[l] = s[l], s[i]
        j = k >> 2
        while j > 0:                                                   # half-cleaners
            for t in range(half_pairs):
                i = ((t & ~(j - 1)) << 1) | (t & (j - 1))
                l = i | j
                if l < n and s[i] > s[l]:
                    s[i], s[l] = s[l], s[i]
            j >>= 1
        k <<= 1
        lg += 1


def test_normalised_bitonic_network_sorts_any_length_in_place():
    rnd = random.Random(3)
    for n in list(range(0, 70)) + [127, 128, 129, 255, 256, 257, 700, 1025]:
        keys = [rnd.randrange(1 << 52) for _ in range(n)]
        want = sorted(keys)
        _segment_sort(keys)
        assert keys == want, n


def _warp_sort_regs(s, e):
    """Index logic of scripts/ubench/seg_sort.cu::warp_sort_regs<E> (32 lanes x E keys in registers; element = lane * E + r;
    strides >= E cross lanes by shuffle; the lower index keeps the minimum; positions >= len hold +inf)."""
    inf = (1 << 64) - 1
    n = len(s)
    v = [[(s[l * e + r] if l * e + r < n else inf) for r in range(e)] for l in range(32)]
    k = 2
    while k <= 32 * e:
        if k <= e:
            for l in range(32):
                for r in range(e):
                    q = r ^ (k - 1)
                    if q > r and v[l][r] > v[l][q]:
                        v[l][r], v[l][q] = v[l][q], v[l][r]
        else:
            mm = k // e - 1
            o = [[v[l ^ mm][e - 1 - r] for r in range(e)] for l in range(32)]
            for l in range(32):
                lower = (l & ((mm + 1) >> 1)) == 0
                for r in range(e):
                    v[l][r] = min(v[l][r], o[l][r]) if lower else max(v[l][r], o[l][r])
        j = k >> 2
        while j > 0:
            if j < e:
                for l in range(32):
                    for r in range(e):
                        if (r & j) == 0 and v[l][r] > v[l][r | j]:
                            v[l][r], v[l][r | j] = v[l][r | j], v[l][r]
            else:
                m = j // e
                v = [[(min(v[l][r], v[l ^ m][r]) if (l & m) == 0 else max(v[l][r], v[l ^ m][r])) for r in range(e)] for l in range(32)]
            j >>= 1
        k <<= 1
    return [v[i // e][i % e] for i in range(n)]


def test_register_shuffle_network_of_the_prepared_ubench():
    rnd = random.Random(5)
    for e in (1, 2, 4, 8):
        for n in {2, 3, 31, 32 * e - 1, 32 * e, 16 * e + 3}:
            keys = [rnd.randrange(1 << 60) for _ in range(min(n, 32 * e))]
            assert _warp_sort_regs(keys, e) == sorted(keys), (e, n)
