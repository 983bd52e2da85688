"""GPU Soft-NMS (aidet_soft_nms_f32) against the reference's OWN soft_nms_cpu_kernel
(mmdet/ops/nms/src/nms_cpu.cpp:70-201), compiled unmodified into oracle/_ref/nms_cpu_ref.so, plus the
reference's known answers (tests/test_soft_nms.py:16-41: 4 boxes -> 4 kept; nms_wrapper.py:81-90 doctest:
6 boxes, sigma 0.5 -> 3)."""
import numpy as np
import pytest
import torch

from aidet_b200.ops import soft_nms
from aidet_b200.ops import functional as F
from oracle import build_ref


def _boxes(n, seed, side=300.0, dup=0.3):
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(n, 2, generator=g) * side
    wh = torch.rand(n, 2, generator=g) * 60 + 4
    b = torch.cat([c - wh / 2, c + wh / 2], 1)
    k = int(n * dup)                                   # near-duplicates so that scores really decay
    b[:k] = b[n - k:] + torch.randn(k, 4, generator=g) * 2
    s = torch.rand(n, generator=g)
    return torch.cat([b, s[:, None]], 1)


@pytest.mark.gpu
def test_soft_nms_reference_known_answers(cuda):
    base = np.array([[49.1, 32.4, 51.0, 35.9, 0.9], [49.3, 32.9, 51.0, 35.3, 0.9], [35.3, 11.5, 39.9, 14.5, 0.4],
                     [35.2, 11.7, 39.7, 15.7, 0.3]])
    for dt in (np.float32, np.float64):                # tests/test_soft_nms.py:22-29
        new, inds = soft_nms(base.astype(dt), 0.7)
        assert new.dtype == dt and inds.dtype == np.int64 and len(inds) == len(new) == 4
    for tt in (torch.FloatTensor, torch.DoubleTensor):  # tests/test_soft_nms.py:31-41
        t = tt(base)
        new, inds = soft_nms(t, 0.7)
        assert new.dtype == t.dtype and inds.dtype == torch.long and len(inds) == 4
    new, inds = soft_nms(torch.tensor(base, dtype=torch.float32, device=cuda), 0.7)
    assert new.is_cuda and inds.is_cuda and len(inds) == 4
    doc = np.array([[4., 3., 5., 3., 0.9], [4., 3., 5., 4., 0.9], [3., 1., 3., 1., 0.5], [3., 1., 3., 1., 0.5],
                    [3., 1., 3., 1., 0.4], [3., 1., 3., 1., 0.0]], dtype=np.float32)
    new, inds = soft_nms(doc, 0.7, sigma=0.5)          # nms_wrapper.py:81-90
    assert len(inds) == len(new) == 3
    with pytest.raises(ValueError):
        soft_nms(doc, 0.7, method='nope')
    e_new, e_inds = soft_nms(np.zeros((0, 5), np.float32), 0.5)
    assert e_new.shape == (0, 5) and e_inds.shape == (0,)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["linear", "gaussian"])
@pytest.mark.parametrize("n", [1, 37, 1500])
def test_soft_nms_matches_reference_kernel(cuda, method, n):
    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref/nms_cpu_ref.so was not built (needs /root/reference at build time)")
    dets = _boxes(n, seed=n)
    code = {"linear": 1, "gaussian": 2}[method]
    want = ref.soft_nms(dets.clone(), 0.3, code, 0.5, 0.05)                   # the reference itself
    new, inds = soft_nms(dets.to(cuda), 0.3, method=method, sigma=0.5, min_score=0.05)
    got = torch.cat([new.cpu(), inds.cpu().float()[:, None]], 1)
    assert got.shape == want.shape
    if method == "linear":                              # same float ops in the same order: bit exact
        assert torch.equal(got, want)
    else:                                               # expf (CUDA) vs std::exp (glibc): <= 2 ulp apart
        assert torch.equal(got[:, 5], want[:, 5]) and torch.equal(got[:, :4], want[:, :4])
        assert torch.allclose(got[:, 4], want[:, 4], rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
def test_soft_nms_batched_groups_match_per_group_calls(cuda):
    ref = build_ref.load()
    dets = _boxes(900, seed=7)
    groups = torch.randint(0, 7, (900,), generator=torch.Generator().manual_seed(1))
    groups[groups == 3] = 4                             # an empty group in the middle
    rows, counts = F.soft_nms_batched(dets.to(cuda), groups.to(cuda), 0.3, 1, 0.5, 0.05, n_groups=7)
    assert counts.tolist()[3] == 0 and int(counts.sum()) == rows.size(0)
    start = 0
    for g in range(7):
        idx = (groups == g).nonzero().flatten()
        part = rows[start:start + int(counts[g])].cpu()
        start += int(counts[g])
        if idx.numel() == 0:
            continue
        if ref is not None:
            want = ref.soft_nms(dets[idx].clone(), 0.3, 1, 0.5, 0.05)
            want[:, 5] = idx[want[:, 5].long()].float()                       # local -> original indices
            assert torch.equal(part, want)
        assert set(part[:, 5].long().tolist()) <= set(idx.tolist())
