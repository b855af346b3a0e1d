"""GPU parity tests (B200): the CUDA path, called through the C ABI, against the float64 oracle on the
same seeded inputs.  Tolerances (BASELINE.json north_star): per-layer outputs within 1e-4 relative on
the fp32 path and 1e-2 on the bf16 path; masks bit exact (tests/test_abi_host.py); MPJPE / P-MPJPE
within 0.1 mm (tests/test_gpu_eval.py)."""
import numpy as np
import pytest
import torch

from oracle import lcn_oracle as O
from tests.gpu_helpers import dev, make_pair, rel_err, synth_xy

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 1e-2}


@pytest.mark.parametrize("mask_type,knn", [("locally_connected", 3), ("locally_connected", 1), ("exponential", 2)])
def test_mask_and_effective_weights(mask_type, knn):
    eng, cfg, p = make_pair(L=1, knn=knn, mask_type=mask_type)
    eng.prepare()
    mask = O.mask_values(cfg, p)
    got = eng.read_tensor(3, 0, 128, 128).cpu().numpy()
    assert rel_err(got, mask) < 2e-6
    for l, name in enumerate(O.weight_names(cfg)):
        wm, _, _ = O.effective_weight(cfg, p[name], mask)
        got = eng.read_tensor(2, l, 128, 128).cpu().numpy()
        assert rel_err(got, wm) < 5e-6, name
        sup = np.kron(cfg.support(), np.ones((wm.shape[0] // 17, wm.shape[1] // 17))) != 0
        assert np.all(got[~sup] == 0), name


def _per_layer_check(eng, cfg, p, x, n, bn_group, rate, tol, keep=None):
    """Compare every Z_l / A_l with the oracle layer applied to the GPU's own previous activation."""
    mask = O.mask_values(cfg, p)
    wn, bn_, bnn = O.weight_names(cfg), O.bias_names(cfg), O.bn_names(cfg)
    n_bn = 1 + 2 * cfg.num_layers
    ng = (n + bn_group - 1) // bn_group
    xpad = np.zeros((ng * bn_group, x.shape[1]))
    xpad[:n] = x
    a_prev = xpad
    worst = 0.0
    for l in range(n_bn):
        wm, _, _ = O.effective_weight(cfg, p[wn[l]], mask)
        z_ref = a_prev @ wm + p[bn_[l]]
        z = eng.read_tensor(0, l, n, bn_group).cpu().numpy().astype(np.float64)
        e = rel_err(z, z_ref)
        worst = max(worst, e)
        assert e < tol, f"Z[{l}] rel err {e}"
        # BN per group on the GPU's Z, then act / dropout / residual
        a_ref = np.empty_like(z)
        for g in range(ng):
            sl = slice(g * bn_group, (g + 1) * bn_group)
            y, _, _, _ = O._bn_forward(z[sl], p[bnn[l] + "/gamma"], p[bnn[l] + "/beta"], cfg.F)
            a_ref[sl] = y
        a_ref = np.where(a_ref > 0, a_ref, 0.2 * a_ref)
        if rate > 0:
            a_ref = a_ref * keep[l] / (1 - rate)
        if l >= 2 and l % 2 == 0 and cfg.residual:
            a_ref = a_ref + eng.read_tensor(1, l - 2, n, bn_group).cpu().numpy()
        a = eng.read_tensor(1, l, n, bn_group).cpu().numpy().astype(np.float64)
        e = rel_err(a, a_ref)
        worst = max(worst, e)
        assert e < tol, f"A[{l}] rel err {e}"
        a_prev = a
    return worst


@pytest.mark.parametrize("path", ["fp32", "bf16"])
@pytest.mark.parametrize("L,knn,n,bn_group", [(1, 3, 200, 200), (2, 2, 300, 128), (1, 3, 256, 256)])
def test_forward_per_layer_and_end_to_end(path, L, knn, n, bn_group):
    eng, cfg, p = make_pair(L=L, knn=knn, path=path)
    x, _ = synth_xy(n)
    out = eng.forward(dev(x), bn_group=bn_group, training=True).cpu().numpy()
    _per_layer_check(eng, cfg, p, x.astype(np.float64), n, bn_group, 0.0, TOL[path])
    ref = O.predict(cfg, p, x.astype(np.float64), bn_group)
    e = rel_err(out, ref)
    assert e < (5e-4 if path == "fp32" else 5e-2), f"end-to-end rel err {e}"
    # inference-layout (rotating buffers) result equals the training-layout one
    out2 = eng.forward(dev(x), bn_group=bn_group, training=False).cpu().numpy()
    if path == "fp32":
        # same kernels, same arithmetic; the two MMA-issuing warps of k_tc_gemm reach the tensor pipe in arrival order,
        # so fp32 accumulation order -- and the last bit -- may differ between two launches
        assert rel_err(out2, out) < 1e-6
    else:   # bf16 inference at bn_group <= 256 runs as the fused cluster kernel (tests/test_gpu_fused.py)
        assert rel_err(out2, out) < 3e-2


@pytest.mark.parametrize("path", ["fp32", "bf16"])
def test_exponential_mask_forward(path):
    eng, cfg, p = make_pair(L=1, knn=1, mask_type="exponential", path=path)
    x, _ = synth_xy(128)
    out = eng.forward(dev(x), bn_group=128, training=True).cpu().numpy()
    _per_layer_check(eng, cfg, p, x.astype(np.float64), 128, 128, 0.0, TOL[path])
    assert rel_err(out, O.forward(cfg, p, x.astype(np.float64))[0]) < (5e-4 if path == "fp32" else 5e-2)


def test_wide_model_F128_forward():
    eng, cfg, p = make_pair(F=128, L=1, knn=2, path="fp32")
    x, _ = synth_xy(128)
    out = eng.forward(dev(x), bn_group=128, training=True).cpu().numpy()
    _per_layer_check(eng, cfg, p, x.astype(np.float64), 128, 128, 0.0, 1e-4)
    assert rel_err(out, O.forward(cfg, p, x.astype(np.float64))[0]) < 5e-4


@pytest.mark.parametrize("mask_type,knn,n", [("locally_connected", 2, 200), ("exponential", 1, 128)])
def test_wide_model_F128_bf16_forward_and_gradients(mask_type, knn, n):
    """F = 128 on the tensor-core path: two 64-channel chunks per joint (34 x 34 chunk grid, up to 34 iterations per CTA
    in the forward / dgrad GEMM, paired half-joint units in the weight-gradient GEMM; the dense exponential mask is the
    largest unit table)."""
    eng, cfg, p = make_pair(F=128, L=1, knn=knn, mask_type=mask_type, path="bf16")
    x, y = synth_xy(n)
    xd, yd = dev(x), dev(y)
    out = eng.forward(xd, bn_group=n, training=True).cpu().numpy()
    _per_layer_check(eng, cfg, p, x.astype(np.float64), n, n, 0.0, TOL["bf16"])
    assert rel_err(out, O.forward(cfg, p, x.astype(np.float64))[0]) < 5e-2
    loss = eng.backward(xd, yd, 0.0).item()
    g = eng.unflatten(eng.true_grads())
    ref_loss, ref_g = O.loss_and_grads(cfg, p, x.astype(np.float64), y.astype(np.float64), 0.0, None)
    assert abs(loss - ref_loss) < 2e-2 * ref_loss
    for k in sorted(ref_g):
        tol = GRAD_TOL["bf16"]
        if k.rsplit("/", 1)[-1].startswith("b") and not k.endswith("b4"):
            tol = 0.25
        if k.endswith("/w1"):
            tol = 0.15     # dZ of 2176 columns rounded to bf16 before the K = batch reduction: 10.5 % of the max measured
        assert rel_err(g[k], ref_g[k]) < tol, f"grad {k}"


def test_predict_zero_pads_last_batch_like_reference():
    eng, cfg, p = make_pair(L=1, knn=3, path="fp32")
    x, _ = synth_xy(300)
    got = eng.predict(x, batch_size=128)
    ref = O.predict(cfg, p, x.astype(np.float64), 128)
    assert got.dtype == np.float64 and got.shape == (300, 51)
    assert rel_err(got, ref) < 5e-4
    # sharding the pose batch at BN-group granularity does not change any result
    got2 = np.concatenate([eng.predict(x[:256], 128), eng.predict(x[256:], 128)])
    assert rel_err(got2, got) < 1e-6            # (last-bit differences only: accumulation in arrival order)


def _keep_masks(eng, cfg, n, rate):
    return [eng.dropout_keep(l, n, rate).cpu().numpy().astype(bool) for l in range(1 + 2 * cfg.num_layers)]


@pytest.mark.parametrize("path", ["fp32", "bf16"])
def test_dropout_forward_with_injected_keep_mask(path):
    eng, cfg, p = make_pair(L=1, knn=3, path=path)
    n, rate = 200, 0.25
    x, _ = synth_xy(n)
    keep = _keep_masks(eng, cfg, n, rate)
    frac = np.mean([k.mean() for k in keep])
    assert abs(frac - 0.75) < 0.01
    out = eng.forward(dev(x), bn_group=n, training=True, dropout=rate).cpu().numpy()
    _per_layer_check(eng, cfg, p, x.astype(np.float64), n, n, rate, TOL[path], keep)
    ref, _ = O.forward(cfg, p, x.astype(np.float64), rate, keep)
    assert rel_err(out, ref) < (5e-4 if path == "fp32" else 5e-2)


# bf16 stores the intermediate gradients (dOut, dZ) in bf16 as tensor-core operands; BN backward subtracts
# means from them, so a few percent of max-norm error on the weight gradients is the expected level
GRAD_TOL = {"fp32": 2e-3, "bf16": 1e-1}


@pytest.mark.parametrize("path", ["fp32", "bf16"])
@pytest.mark.parametrize("L,knn,mask_type,n,rate", [(1, 3, "locally_connected", 200, 0.0),
                                                     (2, 2, "locally_connected", 256, 0.25),
                                                     (1, 1, "exponential", 128, 0.0)])
def test_backward_gradients(path, L, knn, mask_type, n, rate):
    eng, cfg, p = make_pair(L=L, knn=knn, mask_type=mask_type, path=path)
    x, y = synth_xy(n)
    keep = _keep_masks(eng, cfg, n, rate) if rate > 0 else None
    xd, yd = dev(x), dev(y)
    eng.forward(xd, bn_group=n, training=True, dropout=rate)
    loss = eng.backward(xd, yd, rate).item()
    g = eng.unflatten(eng.true_grads())
    ref_loss, ref_g = O.loss_and_grads(cfg, p, x.astype(np.float64), y.astype(np.float64), rate, keep)
    assert abs(loss - ref_loss) < (1e-4 if path == "fp32" else 2e-2) * ref_loss
    assert set(g) == set(ref_g)
    for k in sorted(ref_g):
        e = rel_err(g[k], ref_g[k])
        tol = GRAD_TOL[path]
        # a bias in front of BatchNorm has a gradient that is a near-cancelling sum (BN removes the
        # per-channel mean), so bf16 rounding of dZ shows up amplified there
        if path == "bf16" and k.rsplit("/", 1)[-1].startswith("b") and not k.endswith("b4"):
            tol = 0.25
        assert e < tol, f"grad {k}: rel err {e}"


@pytest.mark.parametrize("path", ["fp32", "bf16"])
def test_train_steps_match_tf1_adam(path):
    eng, cfg, p = make_pair(L=1, knn=3, path=path, perturb=True)
    n = 256
    x, y = synth_xy(n)
    xd, yd = dev(x), dev(y)
    st = O.AdamState()
    p_ref = {k: v.copy() for k, v in p.items()}
    tol = 5e-3 if path == "fp32" else 1e-1
    for step in range(3):
        before = {k: v.copy() for k, v in p_ref.items()}
        ref_loss, ref_lr, _ = O.train_step(cfg, p_ref, st, x.astype(np.float64), y.astype(np.float64))
        loss, lr = eng.train_step(xd, yd)
        assert abs(lr - ref_lr) < 1e-12
        assert abs(loss.item() - ref_loss) < (2e-4 if path == "fp32" else 3e-2) * ref_loss
        got = eng.get_params()
        for k in p_ref:
            d_ref = p_ref[k] - before[k]
            d_got = got[k].astype(np.float64) - before[k]
            scale = np.abs(d_ref).max()
            bad = np.abs(d_got - d_ref) > tol * scale + 1e-7 * np.abs(before[k]).max()
            if path == "fp32":
                assert not bad.any(), k
            else:
                # TF1 Adam moves every element by ~lr*sign(g) once |g| >> eps, so an element whose tiny
                # gradient changes sign under bf16 rounding moves the other way: bound the fraction
                assert bad.mean() < 0.02, (k, bad.mean())
        # keep both sides on the same trajectory: fp32 rounding of the parameters differs slightly
        eng.set_params({k: v.astype(np.float32) for k, v in p_ref.items()})
        for k, (o, r, c) in eng.tensors.items():
            eng.adam_m[o:o + r * c].copy_(torch.as_tensor(st.m[k].astype(np.float32).reshape(-1)))
            eng.adam_v[o:o + r * c].copy_(torch.as_tensor(st.v[k].astype(np.float32).reshape(-1)))
    assert eng.step == 3


def test_library_is_the_loaded_native_code():
    import lcn_pose_b200._lib as L
    maps = open("/proc/self/maps").read()
    assert "liblcn_b200.so" in maps
    assert torch.cuda.get_device_capability(0)[0] == 10
    assert L.load().lcn_version()


def test_graph_replayed_train_steps_equal_eager_steps():
    """CUDA-graph replay of the train step (device-resident step scalars, lcn_step_scalars) reproduces the eager call
    sequence: same Philox dropout counters, same TF1 Adam step sizes -> losses and parameters agree to float
    rounding (the weight-gradient kernel accumulates row splits with floating-point reductions in arrival order)."""
    import torch
    n = 384
    x, y = synth_xy(n)
    xd, yd = dev(x), dev(y)
    for path in ("bf16", "fp32"):
        eng_a, _, _ = make_pair(L=1, knn=2, path=path)
        eng_b, _, _ = make_pair(L=1, knn=2, path=path)
        la, lb = [], []
        for _ in range(4):
            la.append(float(eng_a.train_step(xd, yd, dropout=0.25)[0].item()))
            lb.append(float(eng_b.train_step_graph(xd, yd, dropout=0.25)[0].item()))
        assert np.allclose(la, lb, rtol=1e-5, atol=0), (path, la, lb)     # see docstring: fp reductions in arrival order
        assert eng_a.step == eng_b.step == 4
        # Parameters: equal to float rounding except for isolated elements whose gradient nearly cancels over the batch
        # (|g| ~ 1e-7): there the arrival-order rounding of the weight-gradient reduction moves m / (sqrt(v) + eps) by a
        # visible fraction of one Adam step.  profiles/diag_graph_vs_eager.py measured 0-1 such elements out of 2.46 M
        # for eager-eager, eager-graph and graph-graph pairs alike (max difference 2.7e-6), so a handful is allowed and
        # each is bounded by a small fraction of the 4 x lr = 4e-3 the steps could have moved it.
        d = (eng_a.params - eng_b.params).abs()
        outliers = d > 1e-6 + 1e-4 * eng_b.params.abs()
        assert int(outliers.sum()) <= 8, (path, int(outliers.sum()))
        assert float(d.max()) < 1e-4, (path, float(d.max()))
        dv = (eng_a.adam_v - eng_b.adam_v).abs()
        assert int((dv > 1e-12 + 1e-3 * eng_b.adam_v.abs()).sum()) <= 8


@pytest.mark.parametrize("mask_type,knn,F", [("locally_connected", 3, 64), ("exponential", 1, 64), ("locally_connected", 2, 128)])
def test_packed_gradient_bucket_round_trip(mask_type, knn, F):
    """lcn_model_pack_grads / lcn_model_unpack_grads (the data-parallel exchange buffer): the packed buffer holds
    exactly the nonzero joint-pair blocks of every weight matrix plus the other tensors, and scattering it back
    reproduces every entry backward wrote (the masked-out entries of the bucket are zero and stay untouched)."""
    import torch
    eng, cfg, p = make_pair(F=F, L=1, knn=knn, mask_type=mask_type, path="bf16")
    n = 128
    x, y = synth_xy(n)
    eng.forward(dev(x), bn_group=n, training=True)
    eng.backward(dev(x), dev(y), 0.0)
    raw = eng.grads_raw.clone()
    g0 = eng.unflatten(eng.true_grads())            # per tensor: the alignment gaps of the flat vector are never written
    packed = eng.pack_grads().clone()
    sup = cfg.support() != 0
    want = 0
    for name in O.weight_names(cfg):
        fi, fo = p[name].shape[0] // 17, p[name].shape[1] // 17
        want += int(sup.sum()) * fi * fo
    want += sum(v.size for k, v in p.items() if k not in O.weight_names(cfg))
    assert packed.numel() == want
    assert float(packed.abs().sum()) > 0
    # everything the chain rule / Adam read from the bucket survives the round trip: poison the bucket, scatter the
    # packed copy back, and the true gradients come out bit-identical (entries outside the support -- e.g. the dense
    # edge-layer products -- are never read, so they may keep the poison)
    eng.grads_raw.fill_(12345.0)
    eng.unpack_grads()
    back = eng.grads_raw
    written = back != 12345.0
    assert int(written.sum()) <= want and int(written.sum()) > 0.99 * want      # (an entry may equal the poison by chance)
    assert torch.equal(back[written], raw[written])
    g1 = eng.unflatten(eng.true_grads())
    for k in g0:
        assert np.array_equal(g0[k], g1[k]), k


def test_forward_and_train_step_multi_wave_bf16():
    """More CTAs than SMs in the tensor-core GEMMs (40 row tiles x 4 chunk groups = 160 > 148) and a ragged last tile
    (5000 = 39 * 128 + 8): per-layer parity of the forward pass, and one train step against the oracle's loss."""
    eng, cfg, p = make_pair(L=1, knn=3, path="bf16")
    n = 5000
    x, y = synth_xy(n)
    xd, yd = dev(x), dev(y)
    out = eng.forward(xd, bn_group=n, training=True).cpu().numpy()
    _per_layer_check(eng, cfg, p, x.astype(np.float64), n, n, 0.0, TOL["bf16"])
    assert rel_err(out, O.forward(cfg, p, x.astype(np.float64))[0]) < 5e-2
    loss, _ = eng.train_step(xd, yd)
    ref_loss, _ = O.loss_and_grads(cfg, p, x.astype(np.float64), y.astype(np.float64))
    assert abs(loss.item() - ref_loss) < 2e-2 * ref_loss


@pytest.mark.parametrize("path", ["fp32", "bf16"])
def test_regularization_term_in_gradients_and_adam(path):
    """regularization != 0 (models_att.py:362-365,465-472): reg * sum l2_loss(w*, b*) joins the loss, i.e. reg * theta
    joins the gradient of every w* / b* inside the fused Adam kernel (lcn_model_adam_step's `regularization`)."""
    reg = 5e-4
    eng, cfg, p = make_pair(L=1, knn=3, path=path, reg=reg)
    eng0, _, _ = make_pair(L=1, knn=3, path=path, reg=0.0)
    n = 256
    x, y = synth_xy(n)
    xd, yd = dev(x), dev(y)
    assert abs(eng.l2_regularizer() - O.reg_loss(cfg, p) / reg) < 1e-5 * O.reg_loss(cfg, p) / reg
    # three Adam steps with the regulariser against the oracle (fp32: elementwise; bf16: direction of the step)
    st = O.AdamState()
    p_ref = {k: v.copy() for k, v in p.items()}
    for _ in range(3):
        before = {k: v.copy() for k, v in p_ref.items()}
        ref_loss, _, _ = O.train_step(cfg, p_ref, st, x.astype(np.float64), y.astype(np.float64))
        loss, _ = eng.train_step(xd, yd)
        got = eng.get_params()
        for k in p_ref:
            d_ref, d_got = p_ref[k] - before[k], got[k].astype(np.float64) - before[k]
            bad = np.abs(d_got - d_ref) > (5e-3 if path == "fp32" else 1e-1) * np.abs(d_ref).max() + 1e-7 * np.abs(before[k]).max()
            assert bad.mean() < (1e-9 if path == "fp32" else 0.02), (k, bad.mean())
        eng.set_params({k: v.astype(np.float32) for k, v in p_ref.items()})
        for k, (o, r, c) in eng.tensors.items():
            eng.adam_m[o:o + r * c].copy_(torch.as_tensor(st.m[k].astype(np.float32).reshape(-1)))
            eng.adam_v[o:o + r * c].copy_(torch.as_tensor(st.v[k].astype(np.float32).reshape(-1)))
    # the regulariser's share in isolation: the Adam first moment after one step is 0.1 * (g + reg * theta), so the two
    # engines (same data, same parameters, reg on / off) differ by exactly 0.1 * reg * theta on every w* / b*
    eng0.set_params({k: v.astype(np.float32) for k, v in p.items()})
    eng.set_params({k: v.astype(np.float32) for k, v in p.items()})
    for e in (eng, eng0):
        e.adam_m.zero_(); e.adam_v.zero_(); e.step = 0
        e.train_step(xd, yd)
    m1, m0 = eng.unflatten(eng.adam_m), eng0.unflatten(eng0.adam_m)
    for k in p:
        if k in O.weight_names(cfg) + O.bias_names(cfg):
            assert rel_err(m1[k] - m0[k], 0.1 * reg * p[k]) < (1e-3 if path == "fp32" else 5e-2), k
        elif path == "fp32":
            assert rel_err(m1[k], m0[k]) < 1e-5, k


def test_trainer_state_round_trips_through_a_tf_checkpoint(tmp_path):
    """N steps + save + restore into a fresh engine + N steps == 2N steps: variables, Adam slots and global_step (LR decay,
    bias correction, dropout stream position) all travel through the TensorBundle files (ADVICE: resume)."""
    from lcn_pose_b200.tools import tf_checkpoint
    n = 256
    x, y = synth_xy(n)
    xd, yd = dev(x), dev(y)
    a, _, _ = make_pair(L=1, knn=2, path="fp32")
    b, _, _ = make_pair(L=1, knn=2, path="fp32")
    for _ in range(6):
        a.train_step(xd, yd, dropout=0.25)
    for _ in range(3):
        b.train_step(xd, yd, dropout=0.25)
    prefix = tf_checkpoint.save_model(str(tmp_path), b.step, b.get_params(), b.get_state(), b.tensors)
    c, _, _ = make_pair(L=1, knn=2, path="fp32", seed=9)          # different initial parameters
    params, state = tf_checkpoint.load_model(prefix, c.tensors, c.n_params)
    c.set_params(params)
    c.set_state(state)
    assert c.step == 3
    for _ in range(3):
        c.train_step(xd, yd, dropout=0.25)
    assert c.step == a.step == 6
    d = (a.params - c.params).abs()
    assert float(d.max()) < 1e-5 and int((d > 1e-6 + 1e-4 * a.params.abs()).sum()) <= 8
    assert float((a.adam_v - c.adam_v).abs().max()) <= 1e-3 * float(a.adam_v.abs().max())
