"""Independent torch restatement of the cgcnn graph (network/models_att.py:534-775, 352-421),
written op-for-op like the TF graph and differentiated by torch autograd.  Used to cross-check
the oracle's hand-derived backward, and by bench.py's CPU-baseline leg (dense fp32, all cores:
this is how the reference executes the path -- dense matmuls on masked weights)."""
import torch

J = 17


def build_params(np_params, dtype=torch.float64, requires_grad=True):
    return {k: torch.tensor(v, dtype=dtype, requires_grad=requires_grad) for k, v in np_params.items()}


def forward(cfg, p, x, dropout_rate=0.0, keep_masks=None, names=None):
    from oracle import lcn_oracle as O
    wn, bn_, bnn = O.weight_names(cfg), O.bias_names(cfg), O.bn_names(cfg)
    dtype = x.dtype
    if "exponential" in cfg.mask_type:
        mask = torch.tensor(O.get_exponential_matrix(), dtype=dtype)
    else:
        sup = torch.tensor(cfg.neighbour_matrix.T != 0, dtype=dtype)
        mask = torch.softmax(p["mask"], dim=0) * sup

    def eff(w):
        if cfg.max_norm:
            n = torch.sqrt((w * w).sum())
            w = w * 1.0 / torch.maximum(n, torch.tensor(1.0, dtype=dtype))
        kin, kout = w.shape
        fi, fo = kin // J, kout // J
        return (w.reshape(J, fi, J, fo) * mask.reshape(J, 1, J, 1)).reshape(kin, kout)

    def bn(y, l):
        B = y.shape[0]
        y3 = y.reshape(B, J, cfg.F)
        mu = y3.mean(dim=(0, 1))
        var = ((y3 - mu.detach()) ** 2).mean(dim=(0, 1))      # tf.nn.moments uses stop_gradient(mean)
        inv = torch.rsqrt(var + 1e-3) * p[bnn[l] + "/gamma"]
        return (y3 * inv + (p[bnn[l] + "/beta"] - mu * inv)).reshape(B, J * cfg.F)

    def layer(a, l, act):
        y = a @ eff(p[wn[l]]) + p[bn_[l]]
        if act:
            if cfg.batch_norm:
                y = bn(y, l)
            y = torch.nn.functional.leaky_relu(y, 0.2)
            if dropout_rate > 0:
                y = y * keep_masks[l] / (1 - dropout_rate)
        return y

    y = layer(x, 0, True)
    l = 1
    for _ in range(cfg.num_layers):
        xin = y
        y = layer(layer(xin, l, True), l + 1, True)
        if cfg.residual:
            y = xin + y
        l += 2
    y = layer(y, l, False)
    B = x.shape[0]
    y3 = y.reshape(B, J, 3)
    x3 = x.reshape(B, J, cfg.in_F)
    out = torch.cat([x3[:, :, :2] + y3[:, :, :2], y3[:, :, 2:3]], dim=2)
    return out.reshape(B, J * 3)


def loss_fn(cfg, p, x, labels, dropout_rate=0.0, keep_masks=None):
    from oracle import lcn_oracle as O
    out = forward(cfg, p, x, dropout_rate, keep_masks)
    loss = ((out - labels) ** 2).mean()
    if cfg.regularization is not None and cfg.regularization != 0:
        loss = loss + cfg.regularization * sum((p[n] ** 2).sum() / 2 for n in O.weight_names(cfg) + O.bias_names(cfg))
    return loss, out
