"""GPU parity: mixup / take_log / scaler / frame_shift / add_noise / Adam+EMA / median vs the oracle."""
import random

import numpy as np
import pytest
import torch

from oracle import frontend as ofe, postprocess as opost, trainer as otr
from tests.util import golden, maxdiff

pytestmark = pytest.mark.gpu


def test_mixup_soft_hard_and_rng_order(dev):
    from desed_task_b200.data_augm import mixup
    torch.manual_seed(5)
    data = torch.rand(12, 128, 626)
    tgt = (torch.rand(12, 10, 156) < 0.1).float()
    for kind in ("soft", "hard"):
        np.random.seed(5)
        torch.manual_seed(7)
        c, perm = otr.draw_mixup(12)
        rd, rt = otr.mixup(data, tgt, c, perm, kind)
        np.random.seed(5)
        torch.manual_seed(7)
        md, mt = mixup(data.to(dev), tgt.to(dev), mixup_label_type=kind)
        assert maxdiff(md, rd) < 1e-6 and maxdiff(mt, rt) < 1e-6
    g = golden("augm")
    np.random.seed(5)
    torch.manual_seed(5)
    data = torch.rand(12, 128, 626)
    tgt = (torch.rand(12, 10, 156) < 0.1).float()
    md, mt = mixup(data.to(dev), tgt.to(dev))
    assert np.abs(md[:, :4, :8].cpu().numpy() - g["mixed_head"]).max() < 1e-6
    assert mixup(data.to(dev)).shape == data.shape
    with pytest.raises(NotImplementedError):
        mixup(data.to(dev), tgt.to(dev), mixup_label_type="bogus")


def test_mixup_fused_with_log_and_minmax(dev):
    from desed_task_b200.frontend import take_log, new_minmax, decode_minmax
    torch.manual_seed(0)
    mel = torch.rand(6, 128, 626) * 10
    perm = torch.tensor([3, 0, 5, 1, 2, 4])
    coef = torch.tensor([0.3, 0.3, 0.3, 1.0, 1.0, 1.0])
    ref = mel.clone()
    ref[:3] = 0.3 * mel[:3] + 0.7 * mel[perm[:3]]
    ref = ofe.take_log(ref)
    perm_full = torch.tensor([3, 0, 5, 3, 4, 5])
    mm = new_minmax(6, dev)
    out = take_log(mel.to(dev), perm=perm_full.to(dev), coef=coef.to(dev), minmax=mm)
    assert maxdiff(out, ref) < 1e-4
    mmf = decode_minmax(mm).cpu()
    assert torch.equal(mmf[:, 0], out.amin((1, 2)).cpu()) and torch.equal(mmf[:, 1], out.amax((1, 2)).cpu())


@pytest.mark.parametrize("stat,norm", [("instance", "minmax"), ("instance", "standard"), ("instance", "mean"),
                                       ("dataset", "standard"), ("dataset", "mean"), (None, None)])
def test_scaler_modes(dev, stat, norm):
    from desed_task_b200.utils.scaler import TorchScaler
    torch.manual_seed(1)
    x = torch.randn(4, 128, 626) * 12 + 3
    sc = TorchScaler(stat, norm, (1, 2)) if stat else TorchScaler(None, None)
    kw = {}
    if stat == "dataset":
        loader = [(x[:2],), (x[2:],)]
        sc.fit(loader)
        kw = dict(mean=sc.mean, mean_squared=sc.mean_squared)
    out = sc(x.to(dev))
    ref = ofe.scaler(x, stat, norm, (1, 2), **kw)
    assert maxdiff(out, ref) < 2e-5
    assert sc.state_dict().keys() == ({"mean", "mean_squared"} if stat == "dataset" else set())


def test_scaler_errors():
    from desed_task_b200.utils.scaler import TorchScaler
    with pytest.raises(NotImplementedError):
        TorchScaler("dataset", "minmax")
    with pytest.raises(AssertionError):
        TorchScaler("bogus", "minmax")


def test_frame_shift(dev):
    from desed_task_b200.data_augm import frame_shift
    torch.manual_seed(2)
    mels = torch.rand(12, 128, 626)
    labels = (torch.rand(12, 10, 156) < 0.2).float()
    random.seed(9)
    shifts = otr.draw_frame_shift(12)
    rm, rl = otr.frame_shift(mels, labels, shifts)
    random.seed(9)
    om, ol = frame_shift(mels.to(dev), labels.to(dev))
    assert torch.equal(om.cpu(), rm) and torch.equal(ol.cpu(), rl)
    assert np.array_equal(np.array(shifts), golden("augm")["shifts"])


def test_add_noise(dev):
    from desed_task_b200.data_augm import add_noise
    torch.manual_seed(3)
    mels = torch.rand(5, 128, 626, device=dev)
    torch.manual_seed(11)
    out = add_noise(mels)
    torch.manual_seed(11)
    snr = (6 - 30) * torch.rand((5,), device=dev) + 30
    noise = torch.randn(mels.shape, device=dev)
    ref = otr.add_noise(mels.cpu(), snr.cpu(), noise.cpu())
    assert maxdiff(out, ref) < 1e-5


def test_adam_ema_matches_oracle(dev):
    from desed_task_b200._lib import check, lib, ptr, stream_ptr
    g = torch.Generator().manual_seed(0)
    n = 1112420
    p0 = torch.randn(n, generator=g) * 0.1
    P = {"p": p0.clone()}
    T = {"p": p0.clone()}
    state = {}
    p = p0.clone().to(dev)
    ema = p0.clone().to(dev)
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    for step in range(1, 6):
        gr = torch.randn(n, generator=g) * 0.01
        alpha = otr.update_ema(0.999, step, P, T, ["p"])
        otr.adam_step(P, {"p": gr}, state, ["p"], 1e-3)
        check(lib().sedk_adam_ema(ptr(p), ptr(gr.to(dev)), ptr(m), ptr(v), ptr(ema), n, 1, 1e-3, 0.9, 0.999, 1e-8,
                                  step, alpha, 1.0, stream_ptr()))
    assert maxdiff(p, P["p"]) < 1e-6 and maxdiff(ema, T["p"]) < 1e-6
    # EMA only
    before = p.clone()
    check(lib().sedk_adam_ema(ptr(p), None, None, None, ptr(ema), n, 0, 0, 0, 0, 0, 0, 0.5, 1.0, stream_ptr()))
    assert torch.equal(p, before)


def test_sumsq(dev):
    from desed_task_b200._lib import check, lib, ptr, stream_ptr
    x = torch.randn(100003, device=dev)
    out = torch.zeros(1, dtype=torch.float64, device=dev)
    check(lib().sedk_sumsq(ptr(x), x.numel(), ptr(out), stream_ptr()))
    assert abs(out.item() - (x.double() ** 2).sum().item()) < 1e-6 * out.item()


def test_median_filter(dev):
    from desed_task_b200.utils.postprocess import ClassWiseMedianFilter, median_filter
    g = golden("median")
    sc = torch.from_numpy(g["scores"])                       # [156, 10]
    out = median_filter(sc.t()[None].contiguous().to(dev), 7)        # [1, C, T]
    assert np.array_equal(out[0].t().cpu().numpy(), g["med7"])
    assert np.array_equal(ClassWiseMedianFilter(g["lens"])(g["scores"]), g["med_cw"])
    rng = np.random.RandomState(1)
    big = rng.rand(64, 156, 27).astype(np.float32)
    lens = [int(k) for k in rng.randint(1, 18, 27)]
    got = median_filter(torch.from_numpy(big).to(dev), lens, class_dim=2).cpu().numpy()
    for b in (0, 17, 63):
        assert np.array_equal(got[b], opost.classwise_median_filter(big[b], lens))
    for k in (1, 2, 4, 8, 31):                               # even windows and the maximum size
        got = median_filter(torch.from_numpy(big[:2]).to(dev), k, class_dim=2).cpu().numpy()
        assert np.array_equal(got[1], opost.median_filter_time(big[1], k))


def test_mask_spans_follow_torchaudio_iid_semantics(dev):
    """sedk_mask_spans: value = u*param, start = floor(u'*(size - value)), end = start + floor(value) per example and axis
    (torchaudio mask_along_axis_iid, functional.py:857-869); widths in [0, param), spans inside the axis, fresh draws per
    seed / device counter, param < 1 disables an axis."""
    from desed_task_b200._lib import check, lib, ptr
    B = 4096
    out = torch.zeros(B, 4, dtype=torch.int32, device=dev)
    ctr = torch.zeros(1, dtype=torch.int64, device=dev)
    check(lib().sedk_mask_spans(ptr(out), B, 128, 10, 626, 5, 1234, ptr(ctr), 300, None), "sedk_mask_spans")
    a = out.cpu().numpy().astype(np.int64)
    wf, wt = a[:, 1] - a[:, 0], a[:, 3] - a[:, 2]
    assert (a[:, 0] >= 0).all() and (a[:, 1] <= 128).all() and (wf >= 0).all() and (wf <= 9).all()
    assert (a[:, 2] >= 0).all() and (a[:, 3] <= 626).all() and (wt >= 0).all() and (wt <= 4).all()
    assert abs(wf.mean() - 4.5) < 0.25 and abs(wt.mean() - 2.0) < 0.15          # floor(U[0,10)) / floor(U[0,5))
    assert abs(a[:, 2].mean() - 0.5 * (626 - 2.5)) < 12                         # start ~ U[0, size - value)
    ctr += 1
    out2 = torch.zeros_like(out)
    check(lib().sedk_mask_spans(ptr(out2), B, 128, 10, 626, 5, 1234, ptr(ctr), 300, None), "sedk_mask_spans")
    assert (out2 != out).any()
    check(lib().sedk_mask_spans(ptr(out2), B, 128, 0, 626, 5, 1234, ptr(ctr), 300, None), "sedk_mask_spans")
    assert (out2[:, :2] == 0).all() and (out2[:, 3] >= out2[:, 2]).all()


def test_decode_events_matches_golden(dev):
    """f1: median filter + thresholds + run-length events on the device == the reference's batched_decode_preds numerics
    (fixture minted from the live ManyHotEncoder in oracle/make_golden.py): bit-exact frame boundaries, exact seconds."""
    from desed_task_b200.utils.postprocess import decode_events, frame_to_time, median_filter
    from oracle import postprocess as opost
    g = golden("decode")
    sc = torch.from_numpy(g["scores"]).to(dev)
    ths = [float(t) for t in g["thresholds"]]
    filt = median_filter(sc, 7, class_dim=1)
    assert np.array_equal(filt.transpose(1, 2).cpu().numpy(), g["post"])
    off, ev = decode_events(filt, ths, class_dim=1)
    B, C = sc.shape[0], sc.shape[1]
    got = []
    for ti in range(len(ths)):
        for j in range(B):
            for c in range(C):
                r = (ti * B + j) * C + c
                for on, of in ev[off[r]:off[r + 1]]:
                    got.append((ti, j, c, float(frame_to_time(on)), float(frame_to_time(of))))
    want = [tuple(float(v) if i > 2 else int(v) for i, v in enumerate(row)) for row in g["events"]]
    assert got == want and len(got) > 100
    # ragged clips (pad_indx): the scan stops at each clip's own length; an event still open there ends at that length
    nf = torch.tensor([156, 100, 1, 0], dtype=torch.int32)
    off2, ev2 = decode_events(filt, ths[:2], n_frames=nf, class_dim=1)
    post = g["post"]
    for ti, th in enumerate(ths[:2]):
        for j in range(B):
            for c in range(C):
                r = (ti * B + j) * C + c
                ref = opost.find_contiguous_regions(post[j, :int(nf[j]), c] > np.float32(th))
                assert np.array_equal(ev2[off2[r]:off2[r + 1]].reshape(-1, 2), ref), (ti, j, c)
    # a capacity that is too small is reported and recovered from
    off3, ev3 = decode_events(filt, ths, class_dim=1, capacity=8)
    assert np.array_equal(off3, off) and np.array_equal(ev3, ev)
    # time-major layout [B, T, C] gives the same events
    off4, ev4 = decode_events(filt.transpose(1, 2).contiguous(), ths, class_dim=2)
    assert np.array_equal(off4, off) and np.array_equal(ev4, ev)


def test_batched_decode_preds_mirror(dev):
    """Same arguments / return triple as recipes/dcase2023_task4_baseline/local/utils.py:16-73."""
    from desed_task_b200.utils.postprocess import batched_decode_preds
    from oracle import postprocess as opost
    g = golden("decode")
    sc = torch.from_numpy(g["scores"]).to(dev)

    class Enc:                      # the two members of ManyHotEncoder the function touches
        labels = ["c%d" % i for i in range(10)]

        def _frame_to_time(self, f):
            return opost.frame_to_time(f)
    names = ["/x/clip%d.wav" % i for i in range(sc.shape[0])]
    raw, post, dfs = batched_decode_preds(sc, names, Enc(), thresholds=[0.5, 0.7], median_filter=7)
    assert sorted(raw) == ["clip0", "clip1", "clip2", "clip3"] and list(raw["clip0"].columns[:2]) == ["onset", "offset"]
    assert np.allclose(post["clip1"][Enc.labels].to_numpy(), g["post"][1])
    _, want = opost.batched_decode(g["scores"], Enc.labels, [0.5, 0.7], 7)
    for th in (0.5, 0.7):
        df = dfs[th]
        assert list(df.columns) == ["event_label", "onset", "offset", "filename"]
        rows = [(int(f[4:-4]), l, on, of) for l, on, of, f in df.itertuples(index=False)]
        assert rows == want[th]


def test_encode_strong_batch_matches_golden(dev):
    """f2: batched strong-label encoding on the device == ManyHotEncoder.encode_strong_df (fixture from the live reference):
    list and confidence forms, clipping at both clip ends, overlapping events (later wins), zero-length events, "" labels."""
    import json
    from desed_task_b200.dataio import encode_strong_batch
    g = golden("encode")
    events = json.loads(str(g["events"]))
    labels = ["c%d" % i for i in range(10)]
    y = encode_strong_batch(events, labels, dev)
    assert y.shape == (len(events), 10, 156) and y.dtype == torch.float32
    ref = torch.from_numpy(g["labels"]).float().transpose(1, 2)            # [B, T, C] -> [B, C, T]
    assert torch.equal(y.cpu(), ref)
    # an empty batch entry and a reused output buffer
    y2 = encode_strong_batch([[], events[2]], labels, dev, out=torch.full((2, 10, 156), 7.0, device=dev))
    assert float(y2[0].abs().sum()) == 0.0 and torch.equal(y2[1].cpu(), ref[2])


def test_embedding_storage_format(dev):
    """f3: pre-pooled embeddings.  fp32 pre-pooling reproduces adaptive_avg_pool1d / nearest-exact interpolate (so feeding the
    pooled tensor to the CRNN is the reference's computation: pooling 156 -> 156 frames is the identity); bf16 storage is
    that value rounded to nearest bf16 and `upcast` restores fp32 exactly."""
    import torch.nn.functional as F
    from desed_task_b200.embeddings import pool_embeddings, upcast
    g = torch.Generator().manual_seed(3)
    emb = torch.randn(3, 768, 496, generator=g)
    p32 = pool_embeddings(emb.to(dev), 156, "pool1d", torch.float32)
    ref = F.adaptive_avg_pool1d(emb, 156)
    assert p32.shape == (3, 768, 156) and maxdiff(p32, ref) < 1e-6
    assert maxdiff(pool_embeddings(p32, 156, "pool1d", torch.float32), p32) == 0.0          # identity on pooled input
    pi = pool_embeddings(emb.to(dev), 156, "interpolate", torch.float32)
    assert torch.equal(pi.cpu(), F.interpolate(emb.unsqueeze(1), size=(768, 156), mode="nearest-exact").squeeze(1))
    p16 = pool_embeddings(emb.to(dev), 156, "pool1d", torch.bfloat16)
    assert p16.dtype == torch.bfloat16 and torch.equal(p16.cpu(), p32.cpu().bfloat16())
    assert torch.equal(upcast(p16).cpu(), p16.cpu().float())
    for Te, T in ((496, 156), (100, 156), (157, 156), (1000, 39)):                          # other ratios, incl. upsampling
        e = torch.randn(2, 8, Te, generator=g)
        assert maxdiff(pool_embeddings(e.to(dev), T, "pool1d", torch.float32), F.adaptive_avg_pool1d(e, T)) < 1e-6
        assert torch.equal(pool_embeddings(e.to(dev), T, "interpolate", torch.float32).cpu(),
                           F.interpolate(e.unsqueeze(1), size=(8, T), mode="nearest-exact").squeeze(1))
