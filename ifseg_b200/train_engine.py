"""Training step of the image-free branch on one B200: forward with saved activations + hand-written
backward, as a flat sequence of C-ABI kernel launches (no autograd graph).

Reference call stack: SegCriterion.forward (criterions/seg_criterion.py:165-193) ->
SegOFAModel.forward aux branch (models/segofa/segofa.py:136-151) -> encode_with_artificial_image
(encoder_module.py:499-675) -> surrogate decoder (decoder_module.py:486-677) -> compute_imfree_loss
(seg_criterion.py:246-267) -> optimizer.backward (autograd).  Here:

  * every trainable parameter lives in ONE flat fp32 master buffer (`arena.flat32`; the nn.Parameters are
    views into it, so state_dict / checkpoints / external optimizers keep working) with a parallel flat fp32
    gradient buffer (`param.grad` are views into it: a DDP-style all-reduce is one collective over a
    contiguous range) and flat Adam moments;
  * q/k/v weights of an attention (and the k/v weights of ALL decoder cross-attentions) are adjacent in the
    arena, so the fused [3D,D] / [L*2D,D] GEMM operands and their gradients are plain views;
  * per step every weight matrix is cast to bf16 and transposed once (`sgf_transpose_cast`): W for the forward
    GEMM and dW = dY^T X, W^T for dX = dY W;
  * activations needed by the adjoints are kept (4-5 GB at cfg 3: nothing is recomputed except LayerNorm
    statistics, GELU and the attention probabilities); the residual-stream gradient is one fp32 buffer per
    stack updated in place.

  * the additive position bias is differentiated too: the dQ kernel reduces dS over the batch into a
    per-layer fp32 buffer (vector atomics), `sgf_attn_bias_bwd` scatters it into the relative-position tables and sums
    the layers' abs terms, and the abs term is pulled back through the head-batched pq pk^T GEMM, the position
    projections, the position LayerNorms and the (gathered) position tables.  All 380 tensors the reference gives a
    gradient to receive one; the 20 trainable tensors it never reaches stay without (.grad None), as there.

Round-1 limit: dropout / DropPath are off (deterministic, eval-mode numerics -- the gradient-parity
configuration of SURVEY.md s8d).
"""
from typing import Dict, List, Optional

import torch

from . import ops
from .engine import SegOFAEngine

_BF16 = torch.bfloat16


def _pad8(n):
    return (n + 7) // 8 * 8


class ParamArena:
    """Flat fp32 master / gradient / Adam-moment buffers holding every trainable parameter.

    fp32 model (`alias`): the nn.Parameters and their .grad ARE views of the flat buffers.  Half-precision model
    (`--fp16` / `--bf16`: trainer.py:95-101 calls model.half() / .to(bfloat16) and keeps the fp32 masters in
    FP16Optimizer): the parameters keep their own storage; `pull()` mirrors them into the flat fp32 buffer before a
    forward and gradients are handed back in the parameter dtype (ImFreeBranchFunction)."""

    def __init__(self, groups: List[List[torch.nn.Parameter]], device):
        self.slots: Dict[int, tuple] = {}  # id(param) -> (offset, numel)
        off = 0
        seen = set()
        for grp in groups:
            for p in grp:
                if id(p) in seen:
                    raise ValueError("parameter listed twice in the arena")
                seen.add(id(p))
                self.slots[id(p)] = (off, p.numel())
                off += p.numel()
            off = _pad8(off)
        self.numel = off
        self.flat32 = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad32 = torch.zeros(off, dtype=torch.float32, device=device)
        self.exp_avg = None
        self.exp_avg_sq = None
        self.params = [p for grp in groups for p in grp]
        self.alias = all(p.dtype == torch.float32 for p in self.params)
        self.param_dtype = self.params[0].dtype if self.params else torch.float32
        with torch.no_grad():
            for p in self.params:
                o, n = self.slots[id(p)]
                self.flat32[o:o + n].copy_(p.detach().reshape(-1).float())
                if self.alias:
                    p.data = self.flat32[o:o + n].view(p.shape)
                    p.grad = self.grad32[o:o + n].view(p.shape)

    def has(self, p):
        return p is not None and id(p) in self.slots

    def value(self, p):
        """fp32 master view of parameter p"""
        o, n = self.slots[id(p)]
        return self.flat32[o:o + n].view(p.shape)

    def grad(self, p):
        o, n = self.slots[id(p)]
        return self.grad32[o:o + n].view(p.shape)

    def push(self):
        """mirror mode: fp32 masters -> the parameters (after the engine's own optimizer step; FP16Optimizer's copy-back)"""
        if self.alias:
            return
        with torch.no_grad():
            for p in self.params:
                p.copy_(self.value(p))

    def pull(self):
        """mirror mode: parameter values -> the fp32 master buffer (the caller's optimizer owns the truth)"""
        if self.alias:
            return
        with torch.no_grad():
            dst = [self.value(p) for p in self.params]
            try:
                torch._foreach_copy_(dst, [p.detach() for p in self.params])
            except Exception:  # noqa: BLE001 -- older foreach without dtype conversion
                for d, p in zip(dst, self.params):
                    d.copy_(p.detach())

    def view(self, buf, params, shape):
        """view of `buf` covering the adjacent parameters `params` with the given shape."""
        o0, _ = self.slots[id(params[0])]
        o = o0
        for p in params:
            po, n = self.slots[id(p)]
            if po != o:
                raise ValueError("parameters are not adjacent in the arena")
            o += n
        return buf[o0:o].view(shape)

    def ensure_moments(self):
        if self.exp_avg is None:
            self.exp_avg = torch.zeros_like(self.flat32)
            self.exp_avg_sq = torch.zeros_like(self.flat32)


class ImFreeBranchFunction(torch.autograd.Function):
    """autograd bridge for drop-in callers (optimizer.backward(loss), fp16_optimizer.py:96-106).  The forward runs
    SegOFATrainEngine.forward_train, the backward runs the hand-written adjoint chain.  Every parameter that receives a
    gradient is an INPUT of this node, so whatever hangs on the parameters' gradient accumulators keeps working --
    torch.nn.parallel.DistributedDataParallel's reducer hooks (fairseq's default `pytorch_ddp` backend with
    --find-unused-parameters, distributed_fairseq_model.py:57-83), legacy_ddp's post-backward all-reduce of .grad,
    gradient accumulation over several backward calls (--update-freq).

    The adjoint chain writes into the flat fp32 gradient arena.  What the node returns:
      * fp32 model, parameters without a .grad yet (the usual zero_grad(set_to_none) loop): fresh VIEWS of the arena --
        AccumulateGrad adopts them as param.grad without a copy, so .grad stays inside the flat buffer (one contiguous
        range per all-reduce bucket / fused optimizer);
      * fp32 model, .grad already present (a second backward before the update): the chain accumulates in place and the
        node returns stride-0 zeros, so AccumulateGrad's own `grad += returned` changes nothing but still fires the hooks;
      * half-precision model: one cast of the arena to the parameter dtype, returned as per-parameter views."""

    @staticmethod
    def forward(ctx, engine, aux_input, causal, *params):
        c = engine.forward_train(aux_input, causal=causal)
        ctx.engine, ctx.c = engine, c
        return c["logits"]

    @staticmethod
    def backward(ctx, dlogits):
        eng, c = ctx.engine, ctx.c
        ar = eng.arena
        C = eng.cfg.num_seg
        dl = torch.zeros((c["B"], c["Td"], _pad8(C)), dtype=_BF16, device=dlogits.device)
        dl[..., :C] = dlogits
        in_place = ar.alias and any(p.grad is not None for p in ar.params)
        if in_place:
            eng.attach_grads()
        prev, eng.accumulate = eng.accumulate, in_place
        try:
            eng.backward_from(c, dl)  # zeroes the arena first unless accumulating
        finally:
            eng.accumulate = prev
        ctx.c = None
        if in_place:
            z = torch.zeros((), dtype=torch.float32, device=dlogits.device)
            grads = [z.expand(p.shape) for p in ar.params]
        elif ar.alias:
            grads = [ar.grad(p) for p in ar.params]
        else:
            g16 = ar.grad32.to(ar.param_dtype)
            grads = [g16[o:o + n].view(p.shape) for p in ar.params for (o, n) in (ar.slots[id(p)],)]
        return (None, None, None, *grads)


class _Dense:
    """One GEMM operand set: fp32 master view(s), bf16 W [N,K], bf16 W^T [K,pad8(N)], grads.  `frozen` = the
    nn.Parameters behind a non-arena operand (re-read whenever the model reports an external parameter change)."""
    __slots__ = ("src", "w16", "w16t", "b32", "gw", "gb", "N", "K", "frozen", "frozen_b")


class SegOFATrainEngine:
    _wstream = None  # side stream of the weight-gradient GEMMs (created on first use)
    grad_sync = None

    def __init__(self, model, stochastic=True, seed=1):
        p0 = next(model.parameters())
        if not p0.is_cuda:
            raise RuntimeError("segofa_b200: training runs on a B200 only (no CPU fallback) -- call model.cuda() first")
        dts = {p.dtype for p in model.parameters() if p.requires_grad}
        if len(dts) > 1 or not dts <= {torch.float32, torch.float16, torch.bfloat16}:
            raise RuntimeError(f"segofa_b200 training: trainable parameters must share one floating dtype, got {dts}")
        self.model = model
        self.cfg = model.cfg
        self.device = p0.device
        if self.cfg.decoder_input_type != "encoder_output":
            raise NotImplementedError("training implements --decoder-input-type=encoder_output (the shipped recipe)")
        self._build_arena()
        # the inference engine provides the (batch-invariant) position bias and the real-image no-grad pass
        self.inf: Optional[SegOFAEngine] = None
        self.accumulate = False
        self.grad_sync = None  # callable(lo, hi): gradient range [lo,hi) of arena.grad32 is final (DDP bucket)
        self.step_count = 0
        # training-mode noise of the shipped recipe (--dropout, --encoder/decoder-drop-path-rate; attention and activation
        # dropout are 0 there): counter-based masks regenerated by the adjoint kernels.  stochastic=False gives the
        # deterministic (eval-mode) numerics the gradient-parity tests use.
        a = getattr(model, "args", None)
        self.stochastic = stochastic
        self.seed = seed
        self.drop_p = float(getattr(a, "dropout", 0.0) or 0.0)
        if float(getattr(a, "attention_dropout", 0.0) or 0.0) > 0 or float(getattr(a, "activation_dropout", 0.0) or 0.0) > 0:
            raise NotImplementedError("attention / activation dropout > 0 is not implemented (0 in every shipped recipe)")
        ne, nd = self.cfg.enc_layers, self.cfg.dec_layers
        e_rate = float(getattr(a, "encoder_drop_path_rate", 0.0) or 0.0)
        d_rate = float(getattr(a, "decoder_drop_path_rate", 0.0) or 0.0)
        self.enc_dpr = [e_rate * i / max(ne - 1, 1) for i in range(ne)]  # torch.linspace(0, rate, layers)
        self.dec_dpr = [d_rate * i / max(nd - 1, 1) for i in range(nd)]
        self._fwd_dev = torch.zeros(1, dtype=torch.int32, device=self.device)  # forward counter = mask "step"
        self._fresh = False
        self._wstream = None
        self._hook = None
        self._step_dev = None
        self._scratch: Dict = {}
        self.refresh_weights()

    # ------------------------------------------------------------------------------------
    # parameters
    # ------------------------------------------------------------------------------------
    def _build_arena(self):
        """Arena order = forward order (encoder embeddings, encoder layers, encoder.layer_norm, fused cross k/v,
        decoder embeddings, decoder layers, decoder.layer_norm): the backward finalises gradients from the top of
        the arena downwards, so every all-reduce bucket is one contiguous range (`_sync_down`)."""
        m = self.model
        enc, dec = m.encoder, m.decoder
        groups: List[List[torch.nn.Parameter]] = []
        marks: Dict = {}
        taken = set()
        self._no_grad_names = []
        cross_kv_ids = {id(p) for l in dec.layers for p in (l.encoder_attn.k_proj.weight, l.encoder_attn.v_proj.weight,
                                                            l.encoder_attn.k_proj.bias, l.encoder_attn.v_proj.bias)}

        def add(ps):
            ps = [p for p in ps if p is not None and id(p) not in taken]
            if not ps:
                return
            req = [p.requires_grad for p in ps]
            if not any(req):
                return
            if not all(req):
                raise NotImplementedError("a fused parameter group is only partially trainable")
            groups.append(ps)
            taken.update(id(p) for p in ps)

        def attn(a, cross):
            if cross:
                add([a.q_proj.weight]); add([a.q_proj.bias])
            else:
                add([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight])
                add([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias])
            add([a.out_proj.weight]); add([a.out_proj.bias]); add([a.c_attn])

        def rest(module, prefix, skip_children=()):
            for name, p in module.named_parameters():
                if any(name.startswith(c + ".") for c in skip_children) or id(p) in cross_kv_ids:
                    continue
                if not p.requires_grad or id(p) in taken:
                    continue
                if self._is_bias_path(prefix + name):
                    self._no_grad_names.append(prefix + name)
                    continue
                add([p])

        rest(enc, "encoder.", skip_children=("layers", "layer_norm", "embed_images"))
        for i, l in enumerate(enc.layers):
            marks[("enc", i)] = len(groups)
            attn(l.self_attn, False)
            rest(l, f"encoder.layers.{i}.")
        marks[("enc", len(enc.layers))] = len(groups)
        rest(enc.layer_norm, "encoder.layer_norm.")
        marks["cross_kv"] = len(groups)
        add([w for l in dec.layers for w in (l.encoder_attn.k_proj.weight, l.encoder_attn.v_proj.weight)])
        add([b for l in dec.layers for b in (l.encoder_attn.k_proj.bias, l.encoder_attn.v_proj.bias)])
        rest(dec, "decoder.", skip_children=("layers", "layer_norm"))
        for i, l in enumerate(dec.layers):
            marks[("dec", i)] = len(groups)
            attn(l.self_attn, False)
            attn(l.encoder_attn, True)
            rest(l, f"decoder.layers.{i}.")
        marks[("dec", len(dec.layers))] = len(groups)
        rest(dec.layer_norm, "decoder.layer_norm.")
        self.arena = ParamArena(groups, self.device)
        self.marks = {k: (self.arena.slots[id(groups[gi][0])][0] if gi < len(groups) else self.arena.numel)
                      for k, gi in marks.items()}
        covered = {id(p) for p in self.arena.params}
        for name, p in m.named_parameters():
            if name in self._no_grad_names:
                p.grad = None
            elif p.requires_grad and id(p) not in covered:
                raise RuntimeError(f"trainable parameter {name} is not covered by the training engine")

    def parameters_changed(self):
        """Somebody wrote parameters in place (load_state_dict, lazy seg-token initialisation): the next forward re-casts
        every operand -- frozen ones too -- and the live inference engine (folded BN, stem weights, frozen tables) is
        rebuilt.  CUDA graphs captured before the change replay the old buffers and must be re-captured."""
        self._fresh = False
        self._frozen_stale = True
        self.inf = None

    def live_inference_engine(self):
        """The no-grad engine that shares this engine's operands (rebuilt after parameters_changed())."""
        if self.inf is None:
            self.refresh_weights()
        return self.inf

    def attach_grads(self):
        """param.grad <- views of arena.grad32 (again).  zero_grad(set_to_none=True) drops them: the arena is then
        cleared, which is exactly what the caller asked for."""
        ar = self.arena
        if not ar.alias:
            return
        missing = [p for p in ar.params if p.grad is None]
        if len(missing) == len(ar.params):
            ar.grad32.zero_()
        for p in missing:
            g = ar.grad(p)
            if len(missing) != len(ar.params):
                g.zero_()
            p.grad = g

    def imfree_logits(self, aux_input):
        """logits [B,Td,C] of the image-free branch, attached to autograd (see ImFreeBranchFunction)."""
        return ImFreeBranchFunction.apply(self, aux_input, True, *self.arena.params)

    def real_logits(self, net_input, causal=True):
        """logits [B,Td,C] of the real-image branch, attached to autograd: the supervised loss of
        seg_criterion.py:188-192 (--unsupervised-segmentation=false).  Same node, same adjoint chain: the image rows
        come from the frozen ResNet + image_proj instead of the category-word bags."""
        return ImFreeBranchFunction.apply(self, net_input, causal, *self.arena.params)

    def _wgrad_stream(self):
        if self._wstream is None:
            self._wstream = torch.cuda.Stream()
        return self._wstream

    def _sync_down(self, key):
        """Gradients of arena[marks[key]:] are final: hand the not-yet-synchronised part to the DDP callback."""
        lo = self.marks[key] if key is not None else 0
        if self._wstream is not None and (self.grad_sync is not None or key is None):
            torch.cuda.current_stream().wait_stream(self._wstream)  # the side-stream weight gradients of this range
        if self.grad_sync is not None and lo < self._sync_hi:
            self.grad_sync(lo, self._sync_hi)
        self._sync_hi = min(self._sync_hi, lo)

    @staticmethod
    def _is_bias_path(name):
        """Trainable tensors that no path of the surrogate decoder reaches (the reference leaves their .grad None
        too: SURVEY.md s8a, 20 tensors -- the reason the recipes need --find-unused-parameters)."""
        keys = ("decoder.embed_positions", "decoder.embed_image_positions", "decoder.pos_ln.", "decoder.image_pos_ln.",
                "decoder.token_rel_pos_table_list", "decoder.image_rel_pos_table_list", "code_layernorm_embedding")
        return any(k in name for k in keys)

    def _dense(self, weights, biases):
        """weights/biases: lists of nn.Parameter fused along the output dimension."""
        d = _Dense()
        N = sum(w.shape[0] for w in weights)
        K = weights[0].shape[1]
        d.N, d.K = N, K
        ar = self.arena
        d.frozen, d.frozen_b = None, None
        if ar.has(weights[0]):
            d.src = ar.view(ar.flat32, weights, (N, K))
            d.gw = ar.view(ar.grad32, weights, (N, K))
        else:
            d.frozen = list(weights)
            d.src = torch.cat([w.detach().float() for w in weights], 0).contiguous()
            d.gw = None
        d.w16 = torch.empty((N, K), dtype=_BF16, device=self.device)
        d.w16t = torch.empty((K, _pad8(N)), dtype=_BF16, device=self.device)
        d.b32, d.gb = None, None
        if biases and biases[0] is not None:
            if ar.has(biases[0]):
                d.b32 = ar.view(ar.flat32, biases, (N,))
                d.gb = ar.view(ar.grad32, biases, (N,))
            else:
                d.frozen_b = list(biases)
                d.b32 = torch.cat([b.detach().float() for b in biases], 0).contiguous()
        return d

    def _ln(self, mod):
        ar = self.arena
        if ar.has(mod.weight):
            return (ar.value(mod.weight), ar.value(mod.bias), ar.grad(mod.weight), ar.grad(mod.bias))
        return (mod.weight.detach().float().contiguous(), mod.bias.detach().float().contiguous(), None, None)

    def _vec(self, p):
        if p is None:
            return None, None
        if self.arena.has(p):
            return self.arena.value(p), self.arena.grad(p)
        return p.detach().float().contiguous(), None

    def refresh_weights(self):
        """(Re)derive the bf16 operands from the fp32 masters: once at construction and after every optimizer
        step.  One transpose-cast launch per weight matrix."""
        m = self.model
        enc, dec = m.encoder, m.decoder
        first = not hasattr(self, "dense")
        if first:
            self.dense: List[_Dense] = []
            self.dense_index: Dict[tuple, _Dense] = {}

            def mk(ws, bs):
                d = self._dense(ws, bs)
                self.dense.append(d)
                self.dense_index[tuple(id(w) for w in ws)] = d
                return d

            def attn(a):
                return dict(qkv=mk([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight],
                                   [a.q_proj.bias, a.k_proj.bias, a.v_proj.bias]),
                            out=mk([a.out_proj.weight], [a.out_proj.bias]), c_attn=self._vec(a.c_attn))

            self.enc_layers = []
            for l in enc.layers:
                self.enc_layers.append(dict(
                    attn=attn(l.self_attn), ln_self=self._ln(l.self_attn_layer_norm), ln_attn=self._ln(l.attn_ln),
                    ln_final=self._ln(l.final_layer_norm), ln_ffn=self._ln(l.ffn_layernorm),
                    fc1=mk([l.fc1.weight], [l.fc1.bias]), fc2=mk([l.fc2.weight], [l.fc2.bias])))
            self.cross_kv = mk([w for l in dec.layers for w in (l.encoder_attn.k_proj.weight, l.encoder_attn.v_proj.weight)],
                               [b for l in dec.layers for b in (l.encoder_attn.k_proj.bias, l.encoder_attn.v_proj.bias)])
            self.dec_layers = []
            for l in dec.layers:
                ca = l.encoder_attn
                self.dec_layers.append(dict(
                    attn=attn(l.self_attn),
                    cross=dict(q=mk([ca.q_proj.weight], [ca.q_proj.bias]), out=mk([ca.out_proj.weight], [ca.out_proj.bias]),
                               c_attn=self._vec(ca.c_attn)),
                    ln_self=self._ln(l.self_attn_layer_norm), ln_self_attn=self._ln(l.self_attn_ln),
                    ln_enc_attn=self._ln(l.encoder_attn_layer_norm), ln_cross_attn=self._ln(l.cross_attn_ln),
                    ln_final=self._ln(l.final_layer_norm), ln_ffn=self._ln(l.ffn_layernorm),
                    fc1=mk([l.fc1.weight], [l.fc1.bias]), fc2=mk([l.fc2.weight], [l.fc2.bias])))
            self.seg_proj = mk([dec.seg_projection.weight], [None])
            # additive position bias (encoder_module.py:757-809, decoder_module.py:541-629): projections, LayerNorms,
            # absolute and relative position tables -- all trainable, all rebuilt every step
            self.pos_q, self.pos_k = (mk([m_.weight], [m_.bias]) for m_ in (enc.pos_q_linear, enc.pos_k_linear))
            self.self_pos_q, self.self_pos_k = (mk([m_.weight], [m_.bias]) for m_ in (dec.self_pos_q_linear, dec.self_pos_k_linear))
            self.cross_pos_q, self.cross_pos_k = (mk([m_.weight], [m_.bias]) for m_ in (dec.cross_pos_q_linear, dec.cross_pos_k_linear))
            self.ln_pos, self.ln_img_pos, self.ln_seg_pos = self._ln(enc.pos_ln), self._ln(enc.image_pos_ln), self._ln(dec.seg_pos_ln)
            self.tab_pos, self.tab_img_pos = self._vec(enc.embed_positions.weight), self._vec(enc.embed_image_positions.weight)
            self.tab_seg_pos = self._vec(dec.embed_seg_positions.weight)
            self.rel_tok = [self._vec(t.weight) for t in enc.token_rel_pos_table_list]
            self.rel_img = [self._vec(t.weight) for t in enc.image_rel_pos_table_list]
            self.rel_seg = [self._vec(t.weight) for t in dec.seg_rel_pos_table_list]
            self.token_rp_bucket, self.image_rp_bucket = enc.token_rp_bucket.to(self.device), enc.image_rp_bucket.to(self.device)
            self.seg_rp_bucket = dec.seg_rp_bucket.to(self.device)
            self.ln_emb, self.ln_patch = self._ln(enc.layernorm_embedding), self._ln(enc.patch_layernorm_embedding)
            self.ln_enc_out, self.ln_dec_out = self._ln(enc.layer_norm), self._ln(dec.layer_norm)
            self.dec_ln_emb = self._ln(dec.layernorm_embedding)
            te, self.g_type = self._vec(enc.type_embedding.weight)
            self.type_txt, self.type_img = te[0], te[1]
            self.embed_tokens = enc.embed_tokens.weight.detach()
            self._embed_copy = self.embed_tokens.dtype not in (torch.float32, _BF16) or not self.embed_tokens.is_contiguous()
            if self._embed_copy:  # fp16 table (model.half()): the row kernels read fp32 / bf16
                self.embed_tokens = self.embed_tokens.float().contiguous()
            if enc.embed_tokens.weight.requires_grad:
                raise NotImplementedError("training requires --freeze-encoder-embedding/--freeze-decoder-embedding "
                                          "(the shipped recipe); token-embedding gradients are not implemented")
        stale = getattr(self, "_frozen_stale", False)
        if not first:
            self.arena.pull()  # half-precision model: the caller's optimizer owns the parameters, mirror them
        if stale and self._embed_copy:
            self.embed_tokens = self.model.encoder.embed_tokens.weight.detach().float().contiguous()
        for d in self.dense:
            if stale and d.frozen is not None:  # frozen operand (e.g. the tied seg_projection): re-read the parameters
                d.src.copy_(torch.cat([w.detach().float() for w in d.frozen], 0))
                if d.frozen_b is not None:
                    d.b32.copy_(torch.cat([b.detach().float() for b in d.frozen_b], 0))
            if first or stale or d.gw is not None:
                ops.transpose_cast(d.src, out_t=d.w16t, out_c=d.w16)
        self._frozen_stale = False
        self._fresh = True
        if self.inf is None:
            self.inf = SegOFAEngine(self.model, live=self)
            self.inf.cache_position_bias = False  # the position parameters are trained: never reuse across steps

    # ------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------
    def _buf(self, key, shape, dtype):
        t = self._scratch.get(key)
        if t is None or t.shape != torch.Size(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._scratch[key] = t
        return t

    def _csr(self, name, bucket, ids, lo, row_stride, dtable):
        """(order, offsets, dtable) of one relative-position block: static per shape, cached."""
        key = ("csr", name, ids.numel(), lo, row_stride, dtable.shape[0])
        hit = self._scratch.get(key)
        if hit is None:
            order, flat = ops.bias_block_csr(bucket, ids, lo, row_stride)
            counts = torch.bincount(flat, minlength=dtable.shape[0])
            offsets = torch.zeros(dtable.shape[0] + 1, dtype=torch.int32, device=self.device)
            offsets[1:] = counts.cumsum(0).to(torch.int32)
            hit = (order, offsets)
            self._scratch[key] = hit
        return hit[0], hit[1], dtable

    def _abs_bias(self, pq, pk, Tq, Tk):
        """fp32 [H, Tq, pad64(Tk)] = per-head pq pk^T (one head-batched GEMM)."""
        cfg = self.cfg
        D, H, dh = cfg.embed_dim, cfg.heads, cfg.head_dim
        Tkp = (Tk + 63) // 64 * 64
        out = torch.empty((H, Tq, Tkp), dtype=torch.float32, device=self.device)
        # N = the padded width: key rows >= Tk are TMA zero-fill, so the padding columns come out 0 and the vectorised
        # (N % 32 == 0) epilogue applies
        assert pk.shape[0] >= Tkp  # _key_proj pads the key-side projection with zero rows
        ops.gemm(pq, pk, out, M=Tq, N=Tkp, K=dh, batch=H, lda=D, ldb=D, a_batch_stride=dh, b_batch_stride=dh,
                 ldc=Tkp, c_batch_stride=Tq * Tkp)
        return out

    def _key_proj(self, x, L: _Dense):
        """key-side position projection [Tk, D] inside a zero buffer of pad64(Tk) rows (so that the abs-bias GEMM can run
        at the padded width); returns (padded buffer, [Tk, D] view)."""
        Tk = x.shape[0]
        full = torch.zeros(((Tk + 63) // 64 * 64, L.N), dtype=_BF16, device=self.device)
        ops.gemm(x, L.w16, full[:Tk], bias=L.b32)
        return full, full[:Tk]

    def _position_bias(self, h, w, T_txt):
        """Additive attention biases of both stacks from the CURRENT position parameters, keeping what their adjoint
        needs (post-LN position embeddings and the projected q/k positions)."""
        cfg, dev = self.cfg, self.device
        P, D = h * w, cfg.embed_dim
        T, Td = P + T_txt, P + 1
        oh = cfg.orig_patch_image_size // 16
        if (h, w) != (oh, oh) or (h, w) != (self.model.decoder.seg_bucket_size,) * 2:
            raise NotImplementedError("training on a patch grid different from the orig/seg grid (interpolated position "
                                      "tables) is a 'next' row (SURVEY.md s8f-2)")
        ids = self.inf._image_position_ids(h, w)
        tok_ids = self.inf._cached(("arange", T_txt), lambda: torch.arange(T_txt))
        seg_ids = self.inf._cached(("arange", Td), lambda: torch.arange(Td))
        pos = torch.empty((T, D), dtype=_BF16, device=dev)
        ops.row_layernorm(self.tab_img_pos[0], rows=P, gather_idx=ids, ln2=self.ln_img_pos[:2], out2=pos)
        ops.row_layernorm(self.tab_pos[0], rows=T_txt, ln2=self.ln_pos[:2], out2=pos[P:])
        sc = cfg.pos_scaling
        pq = ops.gemm(pos, self.pos_q.w16, bias=self.pos_q.b32, alpha=sc, alpha_cols=D)
        pk_full, pk = self._key_proj(pos, self.pos_k)
        absb = self._abs_bias(pq, pk_full, T, T)
        enc_biases = []
        for l in range(cfg.enc_layers):
            blocks = [(self.image_rp_bucket, ids, self.rel_img[l][0], 0, P), (self.token_rp_bucket, tok_ids, self.rel_tok[l][0], P, T)]
            enc_biases.append(ops.build_attn_bias(absb, T, blocks, f16=True))
        tgt_pos = torch.empty((Td, D), dtype=_BF16, device=dev)
        ops.row_layernorm(self.tab_seg_pos[0], rows=Td, ln2=self.ln_seg_pos[:2], out2=tgt_pos)
        spq = ops.gemm(tgt_pos, self.self_pos_q.w16, bias=self.self_pos_q.b32, alpha=sc, alpha_cols=D)
        spk_full, spk = self._key_proj(tgt_pos, self.self_pos_k)
        cpq = ops.gemm(tgt_pos, self.cross_pos_q.w16, bias=self.cross_pos_q.b32, alpha=sc, alpha_cols=D)
        cpk_full, cpk = self._key_proj(pos, self.cross_pos_k)
        self_abs = self._abs_bias(spq, spk_full, Td, Td)
        cross_abs = self._abs_bias(cpq, cpk_full, Td, T)
        self_biases = [ops.build_attn_bias(self_abs, Td, [(self.seg_rp_bucket, seg_ids, self.rel_seg[l][0], 0, Td)],
                                           f16=True) for l in range(cfg.dec_layers)]
        cross16 = ops.build_attn_bias(cross_abs, T, (), f16=True)
        # the no-grad real-image pass of the same step (seg_criterion.py:185) sees the same grid and prompt: hand it
        # these biases instead of letting the inference engine rebuild them
        self.inf.cache_position_bias = True
        self.inf._bias_cache = {("enc", h, w, T_txt, False): (list(enc_biases), pos),
                                ("dec", h, w, T): (list(self_biases), cross16)}
        # (fp16: the forward and the adjoint kernels stream the same tensors)
        # key-major copies for the dK/dV kernels (their threads own keys)
        enc_t = [ops.transpose_bias(b_, T) for b_ in enc_biases]
        self_t = [ops.transpose_bias(b_, Td) for b_ in self_biases]
        cross_t = ops.transpose_bias(cross16, T)
        return dict(enc_biases=enc_biases, self_biases=self_biases, cross_abs=cross16,
                    enc_biases_t=enc_t, self_biases_t=self_t, cross_abs_t=cross_t,
                    pos=pos, pq=pq, pk=pk,
                    tgt_pos=tgt_pos, spq=spq, spk=spk, cpq=cpq, cpk=cpk, ids=ids, tok_ids=tok_ids, seg_ids=seg_ids)

    def _abs_bias_bwd(self, dabs, pq, pk, Lq: _Dense, Lk: _Dense, xq, xk, Tq, Tk, dxq, dxk):
        """Adjoint of abs = pq pk^T (per head) and of the two position projections.  dabs fp32 [H,Tq,pad64(Tk)];
        pq/pk bf16 [T*,D] (pq already carries pos_scaling); xq/xk the bf16 projection inputs.  Accumulates the input
        gradients into the fp32 buffers dxq [Tq,D] / dxk [Tk,D] and the parameter gradients into the arena."""
        cfg = self.cfg
        D, H, dh = cfg.embed_dim, cfg.heads, cfg.head_dim
        Tkp = dabs.shape[2]
        d16 = torch.empty((H, Tq, Tkp), dtype=_BF16, device=self.device)
        ops.transpose_cast(dabs.view(H * Tq, Tkp), out_c=d16.view(H * Tq, Tkp), want_t=False)  # fp32 -> bf16
        pkt = ops.transpose_cast(pk)  # [D, pad8(Tk)]
        dpq = torch.empty((Tq, D), dtype=_BF16, device=self.device)
        dpk = torch.empty((Tk, D), dtype=_BF16, device=self.device)
        # d(pos_q W^T + b)_h = pos_scaling * dabs_h pk_h   (one head-batched GEMM)
        ops.gemm(d16, pkt, dpq, M=Tq, N=dh, K=Tk, batch=H, lda=Tkp, ldb=pkt.stride(0), ldc=D, a_batch_stride=Tq * Tkp,
                 b_batch_stride=dh * pkt.stride(0), c_batch_stride=dh, alpha=cfg.pos_scaling, alpha_cols=dh)
        # d(pk)_h = dabs_h^T pq_h: contraction over the query index, both operands read as they lie (MN-major)
        for hh in range(H):
            ops.gemm_ex(d16[hh], pq[:, hh * dh:], dpk[:, hh * dh:], M=Tk, N=dh, K=Tq, a_mn=True, b_mn=True, lda=Tkp,
                        ldb=D, ldc=D, tag="pos_bias")
        self._lin_bwd(dpq, xq, Lq, Tq, "pos_q", dx_out=dxq, dx_residual=dxq)
        self._lin_bwd(dpk, xk, Lk, Tk, "pos_k", dx_out=dxk, dx_residual=dxk)

    def _drop(self, site, rows_per_sample, path_p=0.0):
        if not self.stochastic or (self.drop_p <= 0.0 and path_p <= 0.0):
            return None
        return dict(p=self.drop_p, path_p=path_p, seed=self.seed, site=site, rows_per_sample=rows_per_sample,
                    step=self._fwd_dev)

    def _lin_fwd(self, a, L: _Dense, tag, **kw):
        return ops.gemm(a, L.w16, bias=L.b32, tag=tag, **kw)

    def _lin_bwd(self, dy, x, L: _Dense, M, tag, need_dx=True, dx_dtype=_BF16, dx_out=None, bias_done=False,
                 dx_residual=None):
        """dy bf16 [M,N] (row stride may exceed N), x bf16 [M,K] saved input.  Writes dW / db into the arena
        gradient views, returns dX = dY W."""
        N, K = L.N, L.K
        Mp = _pad8(M)
        if L.gw is not None and K % 32 == 0:
            # Weight / bias gradients feed nothing downstream: they run on a side stream underneath the dX chain (their
            # CTAs fill the SMs that the chain's wave tails and small row kernels leave idle).
            cur = torch.cuda.current_stream()
            ws = self._wgrad_stream()
            ws.wait_stream(cur)
            with torch.cuda.stream(ws):
                if L.gb is not None and not bias_done:
                    ops.transpose_cast(dy, M=M, N=N, want_t=False, colsum=L.gb)  # db = column sums of dY
                # dW[n,k] (+)= sum_t dY[t,n] X[t,k]: both operands are read as they lie (MN-major TMA tiles), the token
                # dimension is split over CTAs and reduced with fp32 atomics into the (zeroed / accumulating) gradient
                tiles = ((N + 127) // 128) * ((K + 127) // 128)
                split = max(1, min(8, (296 + tiles - 1) // tiles, (M + 1023) // 1024))
                ops.gemm_ex(dy, x, L.gw, M=N, N=K, K=M, a_mn=True, b_mn=True, lda=dy.stride(0), ldb=x.stride(0),
                            split_k=split, accumulate=True, tag="wgrad_" + tag)
            dy.record_stream(ws)
            x.record_stream(ws)
        elif L.gw is not None:
            if L.gb is not None and not bias_done:
                ops.transpose_cast(dy, M=M, N=N, want_t=False, colsum=L.gb)
            if True:
                dyt = self._buf(("dyt", Mp), (max(d.N for d in self.dense), Mp), _BF16)
                xt = self._buf(("xt", Mp), (max(d.K for d in self.dense), Mp), _BF16)
                ops.transpose_cast(dy, M=M, N=N, out_t=dyt)
                ops.transpose_cast(x, M=M, N=K, out_t=xt)
                ops.gemm(dyt, xt, L.gw, M=N, N=K, K=M, lda=Mp, ldb=Mp, residual=L.gw, tag="wgrad_" + tag)
        if not need_dx:
            return None
        return ops.gemm(dy, L.w16t, dx_out, M=M, N=K, K=N, lda=dy.stride(0), ldb=L.w16t.stride(0), out_dtype=dx_dtype,
                        residual=dx_residual, tag="dgrad_" + tag)

    # ------------------------------------------------------------------------------------
    # forward + backward of the image-free branch
    # ------------------------------------------------------------------------------------
    def forward_backward(self, aux_input, target_classes, label_smoothing=0.0, grad_scale=1.0, backward=True,
                         check_pads=True):
        """aux_input: the dict segofa.py:136-151 receives (src_tokens, patch_images = bag tokens, patch_masks = bag
        end offsets, prev_output_tokens); target_classes int64 [B,S,S] class ids (<0 or >=C ignored).
        Returns (loss 0-dim tensor = mean pixel CE, logits fp32 [B,Td,C]); gradients of the mean loss times
        grad_scale are left in param.grad (views of arena.grad32)."""
        c = self.forward_train(aux_input, check_pads=check_pads)
        loss, dlogits = self.loss_and_dlogits(c, target_classes, label_smoothing, grad_scale, backward)
        if backward:
            self.backward_from(c, dlogits)
        return loss, c["logits"]

    def loss_and_dlogits(self, c, target_classes, label_smoothing=0.0, grad_scale=1.0, backward=True):
        """compute_imfree_loss on the context of forward_train: (mean pixel CE, bf16 dL/dlogits [B,Td,pad8(C)] or None)."""
        logits, h, w = c["logits"], c["h"], c["w"]
        tgt = target_classes.to(self.device).contiguous()
        pix_lse = torch.empty(tuple(tgt.shape), dtype=torch.float32, device=self.device)
        acc = ops.upsample_ce_loss(logits, tgt, h, w, label_smoothing, lse_out=pix_lse, raw=True)
        loss = acc[0] / acc[1]
        if not backward:
            return loss, None
        dlogits = torch.empty((c["B"], c["Td"], _pad8(self.cfg.num_seg)), dtype=_BF16, device=self.device)
        ops.upsample_ce_loss_bwd(logits, tgt, pix_lse, acc[1:], h, w, dlogits, label_smoothing, grad_scale)
        return loss, dlogits

    def forward_train(self, aux_input, check_pads=True, causal=True):
        """Forward of one training branch keeping what the adjoints need; returns the context dict (`logits` fp32
        [B,Td,C] among it) that backward_from consumes.  `patch_images` decides the branch: integer bag tokens = the
        image-free branch (encoder_module.py:529-551), a float [B,3,S,S] batch = the real-image branch (the frozen
        ResNet stem and image_proj run without gradient, encoder_module.py:191-197).  causal=False is the real branch
        under --full-context-alignment (the image-free branch is always causal, segofa.py:145-149)."""
        if not self._fresh:
            self.refresh_weights()  # an external optimizer may have updated the fp32 masters in place
        self._fresh = False
        self._fwd_dev += 1
        cfg, dev = self.cfg, self.device
        D, H, Fd, C = cfg.embed_dim, cfg.heads, cfg.ffn_dim, cfg.num_seg
        src_tokens = aux_input["src_tokens"].to(dev)
        B, T_txt = src_tokens.shape
        # padded prompts (encoder_module.py:730-742): pad keys are masked in the encoder self-attention and the decoder
        # cross-attention; nothing else reads a pad row, so its value and its gradient never matter.  The check is a
        # host sync (as in the reference, :742); a graph-captured session (check_pads=False) asserts "no pads".
        txt_pad = src_tokens.eq(cfg.padding_idx)
        has_pads = check_pads and bool(txt_pad.any())
        images = aux_input["patch_images"]
        real = images.is_floating_point()
        f32 = torch.float32
        if real:
            if self.arena.has(self.model.encoder.image_proj.weight):
                raise NotImplementedError("training the real-image branch needs --freeze-entire-resnet=true (every shipped "
                                          "recipe): ResNet / image_proj gradients are not implemented")
            pm = aux_input.get("patch_masks")
            if check_pads and pm is not None and not bool(pm.all()):
                raise NotImplementedError("training with masked-out images (patch_masks=False) is not implemented")
            with torch.no_grad():
                feat = self.inf.stem(images.to(dev))  # [B,h,w,1024] bf16, folded FrozenBatchNorm (:170-172)
            h, w = feat.shape[1], feat.shape[2]
        else:
            h = w = cfg.patch_image_size // 16
        self.last_grid = (h, w)
        P = h * w
        T, Td = P + T_txt, P + 1
        M, Md = B * T, B * Td
        kpm = None
        if has_pads:
            kpm = torch.zeros((B, T), dtype=torch.uint8, device=dev)
            kpm[:, P:] = txt_pad
        new = lambda shape, dt=_BF16: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731

        pb = self._position_bias(h, w, T_txt)
        enc_biases, self_biases, cross_abs = pb["enc_biases"], pb["self_biases"], pb["cross_abs"]

        # ------------------------------ encoder forward ------------------------------
        if real:  # image rows = image_proj(ResNet features) (encoder_module.py:416), both frozen
            bag = ops.gemm(feat.view(B * P, 1024), self.inf.w_image_proj, bias=self.inf.b_image_proj, out_dtype=f32)
        else:
            bag = ops.embedding_bag_mean(images.to(dev).contiguous(), aux_input["patch_masks"].to(dev).contiguous(),
                                         self.embed_tokens, P)
        tok_idx = src_tokens.reshape(-1).contiguous()
        x = new((M, D), f32)
        a = new((M, D))
        L0 = self.enc_layers[0]
        ops.row_layernorm(bag, pre_add=self.type_img, ln1=self.ln_patch[:2], out1=x, ln2=L0["ln_self"][:2], out2=a,
                          seg=(P, T, 0), drop=self._drop(1, T))
        ops.row_layernorm(self.embed_tokens, rows=B * T_txt, D=D, gather_idx=tok_idx, pre_add=self.type_txt,
                          ln1=self.ln_emb[:2], out1=x, ln2=L0["ln_self"][:2], out2=a, seg=(T_txt, T, P),
                          drop=self._drop(2, T))
        x_emb = x
        enc_saved = []
        s3 = (3 * D, T * 3 * D)
        for li, L in enumerate(self.enc_layers):
            S = dict(a=a)
            S["qkv"] = self._lin_fwd(a, L["attn"]["qkv"], "qkv", alpha=cfg.attn_scaling, alpha_cols=D)
            S["o"], S["lse"] = new((M, D)), new((B, H, T), f32)
            ops.attention(S["qkv"], S["qkv"][:, D:], S["qkv"][:, 2 * D:], S["o"], B=B, H=H, Tq=T, Tk=T, q_strides=s3,
                          k_strides=s3, v_strides=s3, o_strides=(D, T * D), bias=enc_biases[li],
                          head_scale=L["attn"]["c_attn"][0], key_padding_mask=kpm, lse=S["lse"])
            S["y"] = self._lin_fwd(S["o"], L["attn"]["out"], "out_proj", out_dtype=f32)
            S["x1"], S["a2"] = new((M, D), f32), new((M, D))
            S["dA"], S["dB"] = self._drop(10 + 2 * li, T, self.enc_dpr[li]), self._drop(11 + 2 * li, T, self.enc_dpr[li])
            ops.row_layernorm(S["y"], ln1=L["ln_attn"][:2], residual=x, out1=S["x1"], ln2=L["ln_final"][:2], out2=S["a2"],
                              drop=S["dA"])
            S["h"] = self._lin_fwd(S["a2"], L["fc1"], "fc1")
            S["z"] = new((M, Fd))
            ops.row_layernorm(S["h"], ln2=L["ln_ffn"][:2], out2=S["z"], x_act=ops.ACT_GELU)
            S["x2"] = new((M, D), f32)
            y2 = self._lin_fwd(S["z"], L["fc2"], "fc2", out_dtype=f32)
            nxt = self.enc_layers[li + 1]["ln_self"] if li + 1 < len(self.enc_layers) else self.ln_enc_out
            S["ln_next"] = nxt
            a = new((M, D))
            # x2 = x1 + drop_path(dropout(fc2(...))) and the next block's pre-LayerNorm in one row kernel
            ops.row_layernorm(y2, residual=S["x1"], out1=S["x2"], ln2=nxt[:2], out2=a, drop=S["dB"])
            x = S["x2"]
            enc_saved.append(S)
        enc_out = a  # [B*T, D] bf16

        # ------------------------------ decoder forward ------------------------------
        nL = len(self.dec_layers)
        bos = aux_input["prev_output_tokens"].to(dev)[:, 0].contiguous()
        xd = new((Md, D), f32)
        ad = new((Md, D))
        D0 = self.dec_layers[0]
        ops.row_layernorm(self.embed_tokens, rows=B, D=D, gather_idx=bos, ln1=self.dec_ln_emb[:2], out1=xd,
                          ln2=D0["ln_self"][:2], out2=ad, seg=(1, Td, 0), drop=self._drop(100, Td))
        dec_in_idx = self.inf._cached(("dec_in_idx", B, T, P), lambda: (
            torch.arange(B).unsqueeze(1) * T + torch.arange(P).unsqueeze(0)).reshape(-1))
        ops.row_layernorm(enc_out, rows=B * P, gather_idx=dec_in_idx, ln1=self.dec_ln_emb[:2], out1=xd,
                          ln2=D0["ln_self"][:2], out2=ad, seg=(P, Td, 1), drop=self._drop(101, Td))
        xd_emb = xd
        kv_all = self._lin_fwd(enc_out, self.cross_kv, "cross_kv")  # [M, nL*2D]
        kvs = (nL * 2 * D, T * nL * 2 * D)
        sd3 = (3 * D, Td * 3 * D)
        dec_saved = []
        a, x = ad, xd
        for li, L in enumerate(self.dec_layers):
            S = dict(a=a)
            S["qkv"] = self._lin_fwd(a, L["attn"]["qkv"], "qkv", alpha=cfg.attn_scaling, alpha_cols=D)
            S["o"], S["lse"] = new((Md, D)), new((B, H, Td), f32)
            ops.attention(S["qkv"], S["qkv"][:, D:], S["qkv"][:, 2 * D:], S["o"], B=B, H=H, Tq=Td, Tk=Td, q_strides=sd3,
                          k_strides=sd3, v_strides=sd3, o_strides=(D, Td * D), bias=self_biases[li],
                          head_scale=L["attn"]["c_attn"][0], causal=causal, lse=S["lse"])
            S["y"] = self._lin_fwd(S["o"], L["attn"]["out"], "out_proj", out_dtype=f32)
            S["x1"], S["a2"] = new((Md, D), f32), new((Md, D))
            S["dS"], S["dC"], S["dB"] = (self._drop(110 + 3 * li + k, Td, self.dec_dpr[li]) for k in range(3))
            ops.row_layernorm(S["y"], ln1=L["ln_self_attn"][:2], residual=x, out1=S["x1"], ln2=L["ln_enc_attn"][:2],
                              out2=S["a2"], drop=S["dS"])
            Cx = L["cross"]
            S["qc"] = self._lin_fwd(S["a2"], Cx["q"], "cross_q", alpha=cfg.attn_scaling, alpha_cols=D)
            S["oc"], S["lse_c"] = new((Md, D)), new((B, H, Td), f32)
            kbase = kv_all[:, li * 2 * D:]
            ops.attention(S["qc"], kbase, kbase[:, D:], S["oc"], B=B, H=H, Tq=Td, Tk=T, q_strides=(D, Td * D),
                          k_strides=kvs, v_strides=kvs, o_strides=(D, Td * D), bias=cross_abs,
                          head_scale=Cx["c_attn"][0], key_padding_mask=kpm, lse=S["lse_c"])
            S["yc"] = self._lin_fwd(S["oc"], Cx["out"], "out_proj", out_dtype=f32)
            S["x2"], S["a3"] = new((Md, D), f32), new((Md, D))
            ops.row_layernorm(S["yc"], ln1=L["ln_cross_attn"][:2], residual=S["x1"], out1=S["x2"], ln2=L["ln_final"][:2],
                              out2=S["a3"], drop=S["dC"])
            S["h"] = self._lin_fwd(S["a3"], L["fc1"], "fc1")
            S["z"] = new((Md, Fd))
            ops.row_layernorm(S["h"], ln2=L["ln_ffn"][:2], out2=S["z"], x_act=ops.ACT_GELU)
            S["x3"] = new((Md, D), f32)
            y2 = self._lin_fwd(S["z"], L["fc2"], "fc2", out_dtype=f32)
            nxt = self.dec_layers[li + 1]["ln_self"] if li + 1 < nL else self.ln_dec_out
            S["ln_next"] = nxt
            a = new((Md, D))
            ops.row_layernorm(y2, residual=S["x2"], out1=S["x3"], ln2=nxt[:2], out2=a, drop=S["dB"])
            x = S["x3"]
            dec_saved.append(S)
        feats = a
        logits = ops.gemm(feats, self.seg_proj.w16, out_dtype=f32, tag="seg_proj").view(B, Td, C)

        return dict(logits=logits, h=h, w=w, B=B, T_txt=T_txt, P=P, T=T, Td=Td, M=M, Md=Md, bag=bag, tok_idx=tok_idx,
                    x_emb=x_emb, xd_emb=xd_emb, enc_out=enc_out, kv_all=kv_all, feats=feats, dec_in_idx=dec_in_idx, bos=bos,
                    enc_saved=enc_saved, dec_saved=dec_saved, enc_biases=enc_biases, self_biases=self_biases,
                    cross_abs=cross_abs, pb=pb, causal=causal, real=real, kpm=kpm)

    def backward_from(self, c, dlogits):
        """Adjoint of forward_train: dlogits bf16 [B,Td,pad8(C)] = dL/dlogits.  Parameter gradients are written
        (accumulated when self.accumulate) into arena.grad32, i.e. into param.grad."""
        cfg, dev = self.cfg, self.device
        D, H, Fd, C = cfg.embed_dim, cfg.heads, cfg.ffn_dim, cfg.num_seg
        f32 = torch.float32
        new = lambda shape, dt=_BF16: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
        B, T_txt, P, T, Td, M, Md = (c[k] for k in ("B", "T_txt", "P", "T", "Td", "M", "Md"))
        bag, tok_idx, x_emb, xd_emb, enc_out, kv_all, feats = (c[k] for k in (
            "bag", "tok_idx", "x_emb", "xd_emb", "enc_out", "kv_all", "feats"))
        dec_in_idx, bos, enc_saved, dec_saved = c["dec_in_idx"], c["bos"], c["enc_saved"], c["dec_saved"]
        pb = c["pb"]
        enc_biases, self_biases, cross_abs = pb["enc_biases"], pb["self_biases"], pb["cross_abs"]
        nL = len(self.dec_layers)
        s3, sd3 = (3 * D, T * 3 * D), (3 * D, Td * 3 * D)
        kvs = (nL * 2 * D, T * nL * 2 * D)
        Cp = dlogits.shape[-1]
        if not self.accumulate:
            self.arena.grad32.zero_()
        self._sync_hi = self.arena.numel

        # ------------------------------ decoder backward ------------------------------
        da = self._lin_bwd(dlogits.view(Md, Cp), feats, self.seg_proj, Md, "seg_proj")  # [Md, D] bf16
        dxs = None  # fp32 residual-stream gradient
        dkv_all = new((M, nL * 2 * D))
        pb = c["pb"]
        # gradients w.r.t. the additive position biases: per-layer scratch (consumed + cleared by sgf_attn_bias_bwd),
        # and the layer-summed gradients of the abs terms
        d_self_l = torch.zeros_like(self_biases[0], dtype=f32)
        d_self_abs = torch.zeros_like(self_biases[0], dtype=f32)
        d_cross_abs = torch.zeros_like(cross_abs, dtype=f32)  # no rel term: all layers accumulate straight into it
        for li in reversed(range(nL)):
            L, S = self.dec_layers[li], dec_saved[li]
            nxt = S["ln_next"]
            dx_new = dxs if dxs is not None else new((Md, D), f32)
            dy = new((Md, D))
            ops.row_layernorm_bwd(rows=Md, D=D, v=S["x3"], g2=nxt[0], dy2=da, dv_in=dxs, d_res=dx_new, dx=dy,
                                  dg2=nxt[2], db2=nxt[3], dx_colsum=L["fc2"].gb, drop=S["dB"])
            dxs = dx_new
            dz = self._lin_bwd(dy, S["z"], L["fc2"], Md, "fc2", bias_done=True)
            dh = new((Md, Fd))
            ops.row_layernorm_bwd(rows=Md, D=Fd, x=S["h"], x_act=ops.ACT_GELU, g2=L["ln_ffn"][0], dy2=dz, dx=dh,
                                  dg2=L["ln_ffn"][2], db2=L["ln_ffn"][3], dx_colsum=L["fc1"].gb)
            da3 = self._lin_bwd(dh, S["a3"], L["fc1"], Md, "fc1", bias_done=True)
            # cross-attention block
            dyc = new((Md, D))
            ops.row_layernorm_bwd(rows=Md, D=D, x=S["yc"], g1=L["ln_cross_attn"][0], v=S["x2"], g2=L["ln_final"][0],
                                  dy2=da3, dv_in=dxs, d_res=dxs, dx=dyc, dg1=L["ln_cross_attn"][2],
                                  db1=L["ln_cross_attn"][3], dg2=L["ln_final"][2], db2=L["ln_final"][3],
                                  dx_colsum=L["cross"]["out"].gb, drop=S["dC"])
            Cx = L["cross"]
            doc = self._lin_bwd(dyc, S["oc"], Cx["out"], Md, "out_proj", bias_done=True)
            dqc = new((Md, D))
            kbase, dkbase = kv_all[:, li * 2 * D:], dkv_all[:, li * 2 * D:]
            delta = self._buf("delta", (B, H, max(T, Td)), f32)
            ops.attention_bwd(S["qc"], kbase, kbase[:, D:], S["oc"], doc, dqc, dkbase, dkbase[:, D:], B=B, H=H, Tq=Td,
                              Tk=T, q_strides=(D, Td * D), k_strides=kvs, v_strides=kvs, o_strides=(D, Td * D),
                              do_strides=(D, Td * D), dq_strides=(D, Td * D), dk_strides=kvs, dv_strides=kvs,
                              lse=S["lse_c"], delta=delta, bias=cross_abs, head_scale=Cx["c_attn"][0],
                              d_head_scale=Cx["c_attn"][1], dq_scale=cfg.attn_scaling, dbias=d_cross_abs,
                              bias_t=pb["cross_abs_t"], key_padding_mask=c["kpm"])
            da2 = self._lin_bwd(dqc, S["a2"], Cx["q"], Md, "cross_q")
            # self-attention block
            dy = new((Md, D))
            ops.row_layernorm_bwd(rows=Md, D=D, x=S["y"], g1=L["ln_self_attn"][0], v=S["x1"], g2=L["ln_enc_attn"][0],
                                  dy2=da2, dv_in=dxs, d_res=dxs, dx=dy, dg1=L["ln_self_attn"][2],
                                  db1=L["ln_self_attn"][3], dg2=L["ln_enc_attn"][2], db2=L["ln_enc_attn"][3],
                                  dx_colsum=L["attn"]["out"].gb, drop=S["dS"])
            do = self._lin_bwd(dy, S["o"], L["attn"]["out"], Md, "out_proj", bias_done=True)
            dqkv = new((Md, 3 * D))
            ops.attention_bwd(S["qkv"], S["qkv"][:, D:], S["qkv"][:, 2 * D:], S["o"], do, dqkv, dqkv[:, D:],
                              dqkv[:, 2 * D:], B=B, H=H, Tq=Td, Tk=Td, q_strides=sd3, k_strides=sd3, v_strides=sd3,
                              o_strides=(D, Td * D), do_strides=(D, Td * D), dq_strides=sd3, dk_strides=sd3,
                              dv_strides=sd3, lse=S["lse"], delta=delta, bias=self_biases[li],
                              head_scale=L["attn"]["c_attn"][0], d_head_scale=L["attn"]["c_attn"][1], causal=c["causal"],
                              dq_scale=cfg.attn_scaling, dbias=d_self_l, bias_t=pb["self_biases_t"][li])
            ops.attn_bias_bwd(d_self_l, [self._csr("seg", self.seg_rp_bucket, pb["seg_ids"], 0, d_self_l.stride(1),
                                                   self.rel_seg[li][1])], dabs_acc=d_self_abs)
            da = self._lin_bwd(dqkv, S["a"], L["attn"]["qkv"], Md, "qkv")
            dec_saved[li] = None
            self._sync_down(("dec", li + 1))
        # decoder input embedding: rows 0 = bos (frozen embedding), rows 1..P = encoder_out rows
        d_enc_out = self._lin_bwd(dkv_all, enc_out, self.cross_kv, M, "cross_kv", dx_dtype=f32)  # [M, D] fp32
        g_emb, g_l0 = self.dec_ln_emb, self.dec_layers[0]["ln_self"]
        ops.row_layernorm_bwd(rows=B, D=D, x=self.embed_tokens, gather_idx=bos, g1=g_emb[0], v=xd_emb, g2=g_l0[0],
                              dy2=da, dv_in=dxs, dg1=g_emb[2], db1=g_emb[3], dg2=g_l0[2], db2=g_l0[3], seg=(1, Td, 0),
                              drop=self._drop(100, Td))
        ops.row_layernorm_bwd(rows=B * P, D=D, x=enc_out, gather_idx=dec_in_idx, g1=g_emb[0], v=xd_emb, g2=g_l0[0],
                              dy2=da, dv_in=dxs, dx=d_enc_out, dx_accumulate=True, dg1=g_emb[2], db1=g_emb[3],
                              dg2=g_l0[2], db2=g_l0[3], seg=(P, Td, 1), drop=self._drop(101, Td))

        # decoder position bias: abs terms -> projections -> seg_pos_ln -> embed_seg_positions; the key side of the
        # cross bias reaches the ENCODER's position embeddings (d_pos, continued after the encoder layers)
        d_tgt_pos = torch.zeros((Td, D), dtype=f32, device=dev)
        d_pos = torch.zeros((T, D), dtype=f32, device=dev)
        self._abs_bias_bwd(d_self_abs, pb["spq"], pb["spk"], self.self_pos_q, self.self_pos_k, pb["tgt_pos"], pb["tgt_pos"],
                           Td, Td, d_tgt_pos, d_tgt_pos)
        self._abs_bias_bwd(d_cross_abs, pb["cpq"], pb["cpk"], self.cross_pos_q, self.cross_pos_k, pb["tgt_pos"], pb["pos"],
                           Td, T, d_tgt_pos, d_pos)
        ops.row_layernorm_bwd(rows=Td, D=D, x=self.tab_seg_pos[0], g2=self.ln_seg_pos[0], dy2=d_tgt_pos,
                              dx=self.tab_seg_pos[1], dx_accumulate=True, dg2=self.ln_seg_pos[2], db2=self.ln_seg_pos[3])
        self._sync_down("cross_kv")
        # ------------------------------ encoder backward ------------------------------
        d_enc_l = torch.zeros_like(enc_biases[0], dtype=f32)
        d_enc_abs = torch.zeros_like(enc_biases[0], dtype=f32)
        da = d_enc_out  # fp32 gradient w.r.t. encoder_out (= LN_enc_out(x))
        dxs = None
        for li in reversed(range(len(self.enc_layers))):
            L, S = self.enc_layers[li], enc_saved[li]
            nxt = S["ln_next"]
            dx_new = dxs if dxs is not None else new((M, D), f32)
            dy = new((M, D))
            ops.row_layernorm_bwd(rows=M, D=D, v=S["x2"], g2=nxt[0], dy2=da, dv_in=dxs, d_res=dx_new, dx=dy,
                                  dg2=nxt[2], db2=nxt[3], dx_colsum=L["fc2"].gb, drop=S["dB"])
            dxs = dx_new
            dz = self._lin_bwd(dy, S["z"], L["fc2"], M, "fc2", bias_done=True)
            dh = new((M, Fd))
            ops.row_layernorm_bwd(rows=M, D=Fd, x=S["h"], x_act=ops.ACT_GELU, g2=L["ln_ffn"][0], dy2=dz, dx=dh,
                                  dg2=L["ln_ffn"][2], db2=L["ln_ffn"][3], dx_colsum=L["fc1"].gb)
            da2 = self._lin_bwd(dh, S["a2"], L["fc1"], M, "fc1", bias_done=True)
            dy = new((M, D))
            ops.row_layernorm_bwd(rows=M, D=D, x=S["y"], g1=L["ln_attn"][0], v=S["x1"], g2=L["ln_final"][0], dy2=da2,
                                  dv_in=dxs, d_res=dxs, dx=dy, dg1=L["ln_attn"][2], db1=L["ln_attn"][3],
                                  dg2=L["ln_final"][2], db2=L["ln_final"][3], dx_colsum=L["attn"]["out"].gb,
                                  drop=S["dA"])
            do = self._lin_bwd(dy, S["o"], L["attn"]["out"], M, "out_proj", bias_done=True)
            dqkv = new((M, 3 * D))
            delta = self._buf("delta", (B, H, max(T, Td)), f32)
            ops.attention_bwd(S["qkv"], S["qkv"][:, D:], S["qkv"][:, 2 * D:], S["o"], do, dqkv, dqkv[:, D:],
                              dqkv[:, 2 * D:], B=B, H=H, Tq=T, Tk=T, q_strides=s3, k_strides=s3, v_strides=s3,
                              o_strides=(D, T * D), do_strides=(D, T * D), dq_strides=s3, dk_strides=s3, dv_strides=s3,
                              lse=S["lse"], delta=delta, bias=enc_biases[li], head_scale=L["attn"]["c_attn"][0],
                              d_head_scale=L["attn"]["c_attn"][1], dq_scale=cfg.attn_scaling, dbias=d_enc_l,
                              bias_t=pb["enc_biases_t"][li], key_padding_mask=c["kpm"])
            rs = d_enc_l.stride(1)
            ops.attn_bias_bwd(d_enc_l, [self._csr("img", self.image_rp_bucket, pb["ids"], 0, rs, self.rel_img[li][1]),
                                        self._csr("tok", self.token_rp_bucket, pb["tok_ids"], P, rs, self.rel_tok[li][1])],
                              dabs_acc=d_enc_abs)
            da = self._lin_bwd(dqkv, S["a"], L["attn"]["qkv"], M, "qkv")
            enc_saved[li] = None
            self._sync_down(("enc", li + 1))
        # encoder input embeddings (token embeddings frozen): type embedding + the two embedding LayerNorms
        g_l0 = self.enc_layers[0]["ln_self"]
        gt = self.g_type
        ops.row_layernorm_bwd(rows=B * P, D=D, x=bag, pre_add=self.type_img, g1=self.ln_patch[0], v=x_emb, g2=g_l0[0],
                              dy2=da, dv_in=dxs, dg1=self.ln_patch[2], db1=self.ln_patch[3], dg2=g_l0[2], db2=g_l0[3],
                              d_pre_add=gt[1] if gt is not None else None, seg=(P, T, 0), drop=self._drop(1, T))
        ops.row_layernorm_bwd(rows=B * T_txt, D=D, x=self.embed_tokens, gather_idx=tok_idx, pre_add=self.type_txt,
                              g1=self.ln_emb[0], v=x_emb, g2=g_l0[0], dy2=da, dv_in=dxs, dg1=self.ln_emb[2],
                              db1=self.ln_emb[3], dg2=g_l0[2], db2=g_l0[3], d_pre_add=gt[0] if gt is not None else None,
                              seg=(T_txt, T, P), drop=self._drop(2, T))
        # encoder position bias: abs term -> pos_q/pos_k -> pos_ln / image_pos_ln -> the two position tables
        self._abs_bias_bwd(d_enc_abs, pb["pq"], pb["pk"], self.pos_q, self.pos_k, pb["pos"], pb["pos"], T, T, d_pos, d_pos)
        ops.row_layernorm_bwd(rows=P, D=D, x=self.tab_img_pos[0], gather_idx=pb["ids"], g2=self.ln_img_pos[0], dy2=d_pos,
                              dx=self.tab_img_pos[1], dx_accumulate=True, dg2=self.ln_img_pos[2], db2=self.ln_img_pos[3])
        ops.row_layernorm_bwd(rows=T_txt, D=D, x=self.tab_pos[0], g2=self.ln_pos[0], dy2=d_pos[P:],
                              dx=self.tab_pos[1], dx_accumulate=True, dg2=self.ln_pos[2], db2=self.ln_pos[3])
        self._sync_down(None)

    # ------------------------------------------------------------------------------------
    # optimizer: clip_grad_norm (trainer.py:886) + fused Adam on the flat buffers (fp16_optimizer/adam.py)
    # ------------------------------------------------------------------------------------
    def grad_norm(self):
        out = torch.zeros(1, dtype=torch.float32, device=self.device)
        ops.sumsq(self.arena.grad32, out)
        return out.sqrt()

    def optimizer_step(self, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip_norm=0.0, grad_mult=1.0):
        """grads *= grad_mult; clip to clip_norm (global L2 over the arena); AdamW; refresh the bf16 operands.
        Everything stays on the device (no host sync).  Returns the pre-clip gradient norm (device tensor)."""
        ar = self.arena
        ar.ensure_moments()
        self.step_count += 1
        if self._step_dev is None:
            self._step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._step_dev += 1  # device-resident update counter: the whole step can be replayed from a CUDA graph
        gnorm = self.grad_norm() * grad_mult
        scale = torch.full((1,), float(grad_mult), dtype=torch.float32, device=self.device)
        if clip_norm > 0:
            scale = scale * (clip_norm / (gnorm + 1e-6)).clamp(max=1.0)
        ops.adam_step(ar.flat32, ar.grad32, ar.exp_avg, ar.exp_avg_sq, lr=lr, beta1=betas[0], beta2=betas[1], eps=eps,
                      weight_decay=weight_decay, step=self.step_count, step_dev=self._step_dev, grad_scale=scale)
        ar.push()
        self.refresh_weights()
        return gnorm
