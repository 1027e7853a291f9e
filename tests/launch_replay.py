"""Test infrastructure: replay every kernel launch of a CUDA forward on the CPU from the launch's OWN inputs.

`ReplayTap` plugs into ifseg_b200.ops.set_tap().  Around each GEMM / convolution / attention / row-kernel launch it
snapshots the operands the kernel reads (exactly as they lie in HBM: strided views, fused buffers, in-place residuals),
evaluates the same operation with the rounding-matched oracle's arithmetic (oracle/matched.py: fp32 math on the stored
bf16 / fp16 values, ONE storage rounding at the end) and records the distance to what the kernel wrote.  Because both
sides start from identical inputs there is exactly one storage point between them: this is the granularity at which
north_star's "<=1e-3 rel" is attainable (end to end a quantised chain is chaotic, see oracle/matched.py:Q).
"""
import torch
import torch.nn.functional as F

from oracle import matched as M

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2


def _view(t, shape, strides):
    return torch.as_strided(t, shape, strides).float().cpu()


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


class ReplayTap:
    def __init__(self):
        self.records = []  # dict(kind, tag, shape, rel_l2, ...)

    # ------------------------------------------------------------------ before: snapshot what the kernel will read
    def before(self, kind, **kw):
        snap = dict(kind=kind, tag=kw.get("tag"))
        if kind == "gemm":
            bt, Mm, N, K = kw["batch"], kw["M"], kw["N"], kw["K"]
            snap.update(M=Mm, N=N, K=K, batch=bt, ldc=kw["ldc"], c_bs=kw["c_batch_stride"], act=kw["act"], alpha=kw["alpha"],
                        alpha_cols=kw["alpha_cols"])
            snap["A"] = _view(kw["a"], (bt, Mm, K), (kw["a_batch_stride"], kw["lda"], 1))
            snap["B"] = _view(kw["b"], (bt, N, K), (kw["b_batch_stride"], kw["ldb"], 1))
            for n in ("bias", "scale"):
                snap[n] = kw[n].float().cpu()[:N] if kw[n] is not None else None
            r = kw["residual"]
            snap["R"] = _view(r, (bt, Mm, N), (kw["r_batch_stride"], kw["ldr"], 1)) if r is not None else None
            rn = kw["rownorm"]
            if rn is not None:
                parts = (rn[2] + 63) // 64
                snap["rn"] = (_view(rn[0], (bt, Mm, parts, 2), (Mm * parts * 2, parts * 2, 2, 1)), rn[1].float().cpu(), rn[2])
            snap["stats_out"] = kw["rowstats_out"]
        elif kind == "conv3x3":
            snap.update(x=kw["x"].float().cpu(), w=kw["w"].float().cpu(), scale=kw["scale"].float().cpu(),
                        bias=kw["bias"].float().cpu(), act=kw["act"])
        elif kind == "conv2d":
            snap.update(x=kw["x"].float().cpu(), w=kw["w"].float().cpu(), scale=kw["scale"].float().cpu(),
                        bias=kw["bias"].float().cpu(), act=kw["act"], kh=kw["kh"], kw=kw["kw"], stride=kw["stride"],
                        pad=kw["pad"], window=kw["window"])
        elif kind == "attention":
            B, H, Tq, Tk = kw["B"], kw["H"], kw["Tq"], kw["Tk"]
            snap.update(B=B, H=H, Tq=Tq, Tk=Tk, causal=kw["causal"], o_strides=kw["o_strides"])
            for n, T in (("q", Tq), ("k", Tk), ("v", Tk)):
                rs, bs = kw[n + "_strides"]
                snap[n] = _view(kw[n], (B, H, T, 64), (bs, 64, rs, 1))
            b = kw["bias"]
            snap["bias"] = b.float().cpu()[:, :, :Tk] if b is not None else torch.zeros(H, Tq, Tk)
            hs = kw["head_scale"]
            snap["c_attn"] = hs.float().cpu() if hs is not None else torch.ones(H)
            kp = kw["key_padding_mask"]
            snap["kpm"] = kp.bool().cpu() if kp is not None else None
        elif kind == "row_layernorm":
            assert not kw["drop"], "replay covers the deterministic (inference) configuration"
            rows, D = kw["rows"], kw["D"]
            x = kw["x"]
            if kw["gather_idx"] is not None:
                t = x[kw["gather_idx"][:rows]].float().cpu()[:, :D]
            else:
                t = _view(x, (rows, D), (kw["ldx"], 1))
            snap.update(t=t, rows=rows, D=D, x_act=kw["x_act"])
            snap["pre_add"] = kw["pre_add"].float().cpu() if kw["pre_add"] is not None else None
            snap["ln1"] = tuple(p.float().cpu() for p in kw["ln1"]) if kw["ln1"] else None
            snap["ln2"] = tuple(p.float().cpu() for p in kw["ln2"]) if kw["ln2"] else None
            seg = kw["seg"] or (0, 0, 0)
            r = torch.arange(rows)
            snap["rmap"] = r if seg[0] == 0 else (r // seg[0]) * seg[1] + seg[2] + r % seg[0]
            res = kw["residual"]
            snap["res"] = res[snap["rmap"].to(res.device)].float().cpu()[:, :D] if res is not None else None
            snap["zero"] = kw["zero_row"][:rows].bool().cpu() if kw["zero_row"] is not None else None
        return snap

    # ------------------------------------------------------------------ after: evaluate + compare
    def after(self, s, out):
        kind = s["kind"]
        rec = dict(kind=kind, tag=s.get("tag"))
        if kind == "gemm":
            acc = s["A"] @ s["B"].transpose(1, 2)
            if "rn" in s:
                st, u, dim = s["rn"]
                s0, s1 = st[..., 0].sum(-1, keepdim=True), st[..., 1].sum(-1, keepdim=True)
                mean = s0 / dim
                rstd = torch.rsqrt((s1 / dim - mean * mean).clamp_min(0) + 1e-5)
                acc = rstd * (acc - mean * u[: s["N"]])
            if s["scale"] is not None:
                acc = acc * s["scale"]
            if s["bias"] is not None:
                acc = acc + s["bias"]
            if s["alpha_cols"] > 0:
                acc[..., : s["alpha_cols"]] *= s["alpha"]
            if s["act"] == ACT_GELU:
                acc = F.gelu(acc)
            if s["R"] is not None:
                acc = acc + s["R"]
            if s["act"] == ACT_RELU:
                acc = F.relu(acc)
            got = _view(out, (s["batch"], s["M"], s["N"]), (s["c_bs"], s["ldc"], 1))
            exp = acc.to(out.dtype).float()
            rec.update(shape=(s["batch"], s["M"], s["N"], s["K"]), rel_l2=_rel(got, exp))
            if s["stats_out"] is not None:  # per-row (sum, sumsq) of every 64-column block of the STORED values
                parts = s["N"] // 64
                gs = _view(s["stats_out"], (s["M"], parts, 2), (parts * 2, 2, 1))
                blk = got[0].view(s["M"], parts, 64)
                es = torch.stack([blk.sum(-1), (blk * blk).sum(-1)], dim=-1)
                rec["rowstats_rel_l2"] = _rel(gs, es)
        elif kind == "conv3x3":
            cin = s["x"].shape[-1]
            w4 = s["w"].reshape(s["w"].shape[0], 3, 3, cin)  # [Cout, ky, kx, Cin] (the engine keeps it as a [Cout, 9*Cin] matrix)
            y = F.conv2d(s["x"].permute(0, 3, 1, 2), w4.permute(0, 3, 1, 2), padding=1)
            y = y * s["scale"].view(1, -1, 1, 1) + s["bias"].view(1, -1, 1, 1)
            if s["act"] == ACT_RELU:
                y = F.relu(y)
            exp = y.permute(0, 2, 3, 1).to(torch.bfloat16).float()
            rec.update(shape=tuple(exp.shape), rel_l2=_rel(out.float().cpu(), exp))
        elif kind == "conv2d":
            x, w, cout = s["x"], s["w"], s["w"].shape[0]
            if s["window"] is None:
                cin = x.shape[-1]
                w4 = w.reshape(cout, s["kh"], s["kw"], cin)
                y = F.conv2d(x.permute(0, 3, 1, 2), w4.permute(0, 3, 1, 2), stride=s["stride"], padding=s["pad"])
            else:  # the 7x7/2 stem convolution: the kernel read a zero-padded 8-channel image as 8-pixel windows
                H, W, pad, cin, kw_real = s["window"]
                xl = x[:, pad:pad + H, pad:pad + W, :cin]
                w4 = w.reshape(cout, s["kh"], 8, 8)[:, :, :kw_real, :cin]
                assert w.reshape(cout, s["kh"], 8, 8)[:, :, kw_real:].abs().max() == 0 and x[..., cin:].abs().max() == 0
                y = F.conv2d(xl.permute(0, 3, 1, 2), w4.permute(0, 3, 1, 2), stride=s["stride"], padding=pad)
            y = y * s["scale"].view(1, -1, 1, 1) + s["bias"].view(1, -1, 1, 1)
            if s["act"] == ACT_RELU:
                y = F.relu(y)
            exp = y.permute(0, 2, 3, 1).to(torch.bfloat16).float()
            rec.update(shape=tuple(exp.shape), rel_l2=_rel(out.float().cpu(), exp))
        elif kind == "attention":
            exp = M.attention(s["q"], s["k"], s["v"], s["bias"], s["c_attn"], M.Q(True), bool(s["causal"]), s["kpm"])
            ors, obs = s["o_strides"]
            got = _view(out, (s["B"], s["Tq"], s["H"] * 64), (obs, ors, 1))
            rec.update(shape=(s["B"], s["H"], s["Tq"], s["Tk"]), rel_l2=_rel(got, exp))
        elif kind == "row_layernorm":
            t, D = s["t"], s["D"]
            if s["x_act"] == ACT_GELU:
                t = F.gelu(t)
            if s["pre_add"] is not None:
                t = t + s["pre_add"]
            u = F.layer_norm(t, (D,), s["ln1"][0], s["ln1"][1], 1e-5) if s["ln1"] else t
            v = u + s["res"] if s["res"] is not None else u
            if s["zero"] is not None:
                v = torch.where(s["zero"].unsqueeze(1), torch.zeros_like(v), v)
            out1, out2 = out
            errs = []
            if out1 is not None:
                got = out1[s["rmap"].to(out1.device)].float().cpu()[:, :D]
                errs.append(_rel(got, v.to(out1.dtype).float()))
            if out2 is not None:
                got = out2[s["rmap"].to(out2.device)].float().cpu()[:, :D]
                exp = F.layer_norm(v, (D,), s["ln2"][0], s["ln2"][1], 1e-5).to(torch.bfloat16).float()
                errs.append(_rel(got, exp))
            rec.update(shape=(s["rows"], D), rel_l2=max(errs))
        self.records.append(rec)

    def summary(self):
        out = {}
        for r in self.records:
            d = out.setdefault(r["kind"], dict(launches=0, max_rel_l2=0.0, worst=None))
            d["launches"] += 1
            if r["rel_l2"] > d["max_rel_l2"]:
                d["max_rel_l2"], d["worst"] = r["rel_l2"], f"{r.get('tag')} {r['shape']}"
            if "rowstats_rel_l2" in r:
                d["max_rowstats_rel_l2"] = max(d.get("max_rowstats_rel_l2", 0.0), r["rowstats_rel_l2"])
        return out
