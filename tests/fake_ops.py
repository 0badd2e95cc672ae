"""A torch-CPU stand-in for pantomatrix_b200.ops, for HOST-LOGIC tests only (tests/test_host_logic.py).

The build container has no GPU.  To exercise the Python scheduling code of the product (window plan,
audio hoisting, weight packing / BatchNorm folding, seed decode, strides and views) without a device, the
kernel wrappers are replaced by these restatements of each kernel's contract (include/pm_emage.h).  This
file lives under tests/ and is never importable from the product package."""
import torch
import torch.nn.functional as F

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
ROW_NONE, ROW_PE, ROW_SPK = 0, 1, 2
launch_count = 0
PLANE_DTYPE = None                # None: follow ops.set_plane_format(); tests/diag_split_formats.py overrides it

import pantomatrix_b200.ops as _real  # noqa: E402
from pantomatrix_b200.ops import Act, PackedW, Planes, _round_up  # noqa: E402,F401  (same containers as the product)


def _plane_dtype():
    return PLANE_DTYPE or _real._PLANE_DTYPE


def _split(v, nsplit):
    planes, rem = [], v
    for _ in range(nsplit):
        p = rem.to(_plane_dtype())
        planes.append(p)
        rem = rem - p.float()
    return planes


def _mk_planes(y, nsplit, slack_rows=0):
    """(batch, rows, ch) fp32 -> Planes with NaN-poisoned padding (catches reads of uninitialised memory)."""
    batch, rows, ch = y.shape
    ld = _round_up(ch, 8)
    buf = torch.full((nsplit, batch * rows + slack_rows, ld), float("nan"), dtype=_plane_dtype())
    buf[:, batch * rows:] = 0
    pre = _real.F16_ACT_SCALE if _plane_dtype() == torch.float16 else 1.0      # same exact pre-scale as the kernels
    for i, p in enumerate(_split(y.reshape(batch * rows, ch).float() * pre, nsplit)):
        buf[i, :batch * rows, :ch] = p
    return Planes(buf[:, :batch * rows].view(nsplit, batch, rows, ld), rows, ch, slack_rows)


def _res(y, nsplit, f32, lead=None):
    if nsplit == 0:
        return y
    y3 = y if lead is None else y.reshape(*lead, y.shape[-1])
    return Act(y if f32 else None, _mk_planes(y3, nsplit))


def _act(v, act, slope):
    return F.relu(v) if act == ACT_RELU else F.leaky_relu(v, slope) if act == ACT_LEAKY else v


def tapgemm(a, w, bias, *, rows_out=None, stride=1, pad=0, act=ACT_NONE, slope=0.0, residual=None, out=None):
    y = F.conv1d(a.transpose(1, 2), w.permute(1, 2, 0), bias, stride=stride, padding=pad).transpose(1, 2)
    if rows_out is not None:
        y = y[:, :rows_out]
    if residual is not None:
        y = y + residual
    y = _act(y, act, slope)
    if out is not None:
        out.copy_(y)
        return out
    return y.contiguous()


def wav_stem(audio, a_bs, a_ws, batch, windows, n_samples, w1, b1, wd, bd, *, stride, pad, slope, offset=0, nsplit=0):
    flat = audio.reshape(-1)
    seqs = torch.stack([flat[offset + b * a_bs + w * a_ws: offset + b * a_bs + w * a_ws + n_samples]
                        for w in range(windows) for b in range(batch)])            # window-major
    x = seqs.unsqueeze(1)
    y1 = F.leaky_relu(F.conv1d(x, w1.unsqueeze(1), b1, stride=stride, padding=pad), slope).transpose(1, 2)
    sc = F.conv1d(x, wd.unsqueeze(1), bd, stride=stride, padding=pad).transpose(1, 2)
    y1 = y1.contiguous()
    return (_mk_planes(y1, nsplit) if nsplit else y1), sc.contiguous()


def add_layernorm(x, r, gamma, beta, eps=1e-5, nsplit=0, f32=True):
    return _res(F.layer_norm(x if r is None else x + r, (x.shape[-1],), gamma, beta, eps), nsplit, f32)


def attention(q, k, v, batch, heads, tq, tk, head_dim, nsplit=0, f32=True):
    E = heads * head_dim
    qq = q[:, :E].reshape(batch, tq, heads, head_dim).transpose(1, 2)
    kk = k[:, :E].reshape(batch, tk, heads, head_dim).transpose(1, 2)
    vv = v[:, :E].reshape(batch, tk, heads, head_dim).transpose(1, 2)
    att = torch.softmax(qq @ kk.transpose(-1, -2) / head_dim ** 0.5, -1)
    return _res((att @ vv).transpose(1, 2).reshape(batch * tq, E), nsplit, f32, lead=(batch, tq))


def attention_tc(q, q_col0, k, k_col0, v, v_col0, batch, heads, tq, tk, head_dim, nsplit=2, f32=False):
    E = heads * head_dim
    val = lambda pl, c0, rows: (pl.t[:, :, :rows, c0:c0 + E].float().sum(0) / _real.F16_ACT_SCALE).reshape(batch * rows, E)
    return attention(val(q, q_col0, tq), val(k, k_col0, tk), val(v, v_col0, tk), batch, heads, tq, tk, head_dim, nsplit=nsplit, f32=f32)


def add_rows(x, pe, spk, first, second, batch, rows, ch, nsplit=0, f32=True):
    v = torch.zeros(batch, rows, ch) if x is None else x.reshape(batch, rows, ch)
    for code in (first, second):
        if code == ROW_PE:
            v = v + pe[None, :rows]
        elif code == ROW_SPK:
            v = v + spk[:, None]
    return _res(v.contiguous(), nsplit, f32)


def add2(a, b, nsplit=0, f32=True):
    return _res(a + b, nsplit, f32)


def window_input(motion, mask, seed, mask_embedding, start, win_len, pre, nsplit=0, f32=True, shape=None):
    if motion is None:                                   # inference()'s defaults (M.py:369-377)
        motion = torch.zeros(shape)
        motion[:, :, 0:shape[2] - 7:6] = 1.0
        motion[:, :, 4:shape[2] - 7:6] = 1.0
    if mask is None:
        mask = torch.ones(motion.shape)
    wm, wk = motion[:, start:start + win_len].clone(), mask[:, start:start + win_len].clone()
    if pre:
        if seed is not None:
            wm[:, :pre] = torch.where(wk[:, :pre] == 0, motion[:, start:start + pre], seed)
        wk[:, :pre] = 0
    return _res(torch.where(wk == 1, mask_embedding.view(1, 1, -1).expand_as(wm), wm), nsplit, f32)


def l2_argmin(z, codebook, e2, engine="auto", max_ctas=0):
    flat = z.reshape(-1, codebook.shape[1])
    d = (flat ** 2).sum(1, keepdim=True) + e2 - 2 * flat @ codebook.t()
    return d.argmin(1).reshape(z.shape[:-1])


def row_argmax(x, nonfinite=None):
    if nonfinite is not None and not bool(torch.isfinite(x).all()):
        nonfinite.fill_(1)
    return x.argmax(-1)


def zero_flag(device):
    return torch.zeros(1, dtype=torch.int32)


def gather_rows(codebook, index, nsplit=0, f32=True):
    y = codebook[index]
    return _res(y, nsplit, f32, lead=(index.shape[0], index.numel() // index.shape[0]) if index.dim() > 1 else (1, index.numel()))


def row_sqnorm(x):
    return (x ** 2).sum(1)


def pose_compose(face, upper, hands, lower, bs, t, device):
    from oracle import emage_oracle as O
    z = lambda n: torch.zeros(bs, t, n)
    aa6 = lambda x: O.rot6d_to_axis_angle(x.reshape(bs, t, -1, 6)).reshape(bs, t, -1)
    jaw = O.rot6d_to_axis_angle(face[:, :, :6]) if face is not None else z(3)
    expr = face[:, :, 6:] if face is not None else z(100)
    up = aa6(upper) if upper is not None else z(39)
    ha = aa6(hands) if hands is not None else z(90)
    lo = aa6(lower[:, :, :54]) if lower is not None else z(27)
    tf = lower[:, :, 54:] if lower is not None else z(7)
    aa = (O._scatter_joints(up, O.UPPER_JOINTS, bs, t) + O._scatter_joints(ha, O.HANDS_JOINTS, bs, t)
          + O._scatter_joints(lo, O.LOWER_JOINTS, bs, t))
    aa[:, :, 66:69] = jaw
    m4 = torch.cat([O.axis_angle_to_rot6d(aa.reshape(bs, t, 55, 3)).reshape(bs, t, 330), tf], 2)
    return expr.contiguous(), aa, m4


def global_trans(rec, ref_trans, dt, vel_off=54):
    v = rec[:, :, vel_off:vel_off + 3]
    x, z = [ref_trans[:, 0:1]], [ref_trans[:, 2:3]]
    for i in range(1, rec.shape[1]):
        x.append(v[:, i - 1, 0:1] * dt + x[-1])
        z.append(v[:, i - 1, 2:3] * dt + z[-1])
    return torch.stack([torch.cat(x, 1), v[:, :, 1], torch.cat(z, 1)], -1)


# ---- tensor-core engine stand-ins (same Planes / PackedW containers as the product) --------------------


def split_bf16(x, nsplit, slack_rows=0):
    return _mk_planes(x, nsplit, slack_rows)


def tapgemm_tc(a, w, bias, *, rows_in=None, rows_out, pad=0, act=ACT_NONE, act_cols=0, slope=0.0, residual=None,
               want_f32=True, out_nsplit=0, out=None, a_view=None, out_slack=0, prefetch=None):
    t = a.t
    nsplit, batch = t.shape[0], t.shape[1]
    rows_a, cin, lda = (a.rows, a.ch, t.stride(2)) if a_view is None else a_view
    # the logical (batch, rows_a, cin) view of the plane memory, exactly as the TMA descriptor addresses it
    x = sum(torch.as_strided(t[i], (batch, rows_a, cin), (t.stride(1), lda, 1)).float() for i in range(nsplit))
    wf = w.t[:, :, :w.cout, :w.cin].float().sum(0)                       # (taps, cout, cin)
    # products the kernel forms: all (i, j) with i + j < nsplit
    y = 0
    for i in range(nsplit):
        xi = torch.as_strided(t[i], (batch, rows_a, cin), (t.stride(1), lda, 1)).float()
        for j in range(nsplit - i):
            wj = w.t[j, :, :w.cout, :w.cin].float()
            y = y + F.conv1d(xi.transpose(1, 2), wj.permute(1, 2, 0), None, padding=pad).transpose(1, 2)
    y = y[:, :rows_out] * getattr(w, "acc_scale", 1.0)          # fp16 weights are packed pre-scaled by a power of two
    if bias is not None:
        y = y + bias
    if residual is not None:
        y = y + residual
    if act != ACT_NONE:
        cols = w.cout if act_cols <= 0 else act_cols
        y = torch.cat([_act(y[..., :cols], act, slope), y[..., cols:]], -1)
    assert torch.isfinite(y).all(), "read uninitialised plane memory"
    out_p = _mk_planes(y, out_nsplit, out_slack) if out_nsplit else None
    if want_f32:
        if out is not None:
            out.copy_(y)
            y = out
        return y.contiguous() if out is None else y, out_p
    return None, out_p


# ---- CaMN / DisCo stand-ins ---------------------------------------------------------------------------


def lstm_bidir(xproj, whh, barrier, hidden):
    bs, T, _ = xproj.shape
    outs = []
    for d in range(2):
        xp, w = xproj[:, :, d * 4 * hidden:(d + 1) * 4 * hidden], whh[d]
        h, c = torch.zeros(bs, hidden), torch.zeros(bs, hidden)
        seq = [None] * T
        for t in (range(T) if d == 0 else range(T - 1, -1, -1)):
            i, f, g, o = (xp[:, t] + h @ w.t()).split(hidden, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            seq[t] = h
        outs.append(torch.stack(seq, 1))
    return torch.cat(outs, 2)


def rot6d_to_aa(rot6d, slot, n_sel):
    from oracle import emage_oracle as O
    lead = rot6d.shape[:-1]
    aa = O.rot6d_to_axis_angle(rot6d.reshape(-1, n_sel, 6))
    full = torch.zeros(aa.shape[0], 55, 3)
    for j in range(55):
        if int(slot[j]) >= 0:
            full[:, j] = aa[:, int(slot[j])]
    return full.reshape(*lead, 165)


def softmax2_mix(sel, c1, c2, out=None):
    w = torch.softmax(sel, dim=-1)
    y = w[..., 0:1] * c1 + w[..., 1:2] * c2
    if out is not None:
        out.copy_(y)
        return out
    return y
