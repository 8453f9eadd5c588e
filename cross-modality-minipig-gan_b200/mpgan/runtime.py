"""Per-network device runtime: flat fp32 master / gradient / Adam buffers with the nn.Parameters re-pointed
at views of them (conv weights in the kernels' OTI layout, exposed through permuted views so shapes and
state-dict keys stay the reference's), plus the bf16 shadows the tcgen05 kernels read.

The flat layout is what makes the optimiser one launch (torch.optim.Adam at GAN_final.py:298-308 loops over
450 + 18 tensors) and the data-parallel gradient exchange one NCCL call per network.
"""
import torch
import torch.nn as nn

from . import ops

_CONV_TYPES = (nn.Conv2d, nn.Conv3d, nn.ConvTranspose2d, nn.ConvTranspose3d)
_ALIGN = 8  # elements: 32 B in fp32, 16 B in the bf16 shadow (TMA base alignment)


class ConvRec:
    """Everything the kernels need for one Conv / ConvTranspose module."""
    __slots__ = ("spec", "w32", "w", "wt", "bias", "dw", "db", "need_wt")


class Runtime:
    def __init__(self, net, precision):
        assert precision in ("bf16", "fp32")
        self.net = net
        self.precision = precision
        self.dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.device = None
        self.flat = self.grad = self.exp_avg = self.exp_avg_sq = self.shadow = None
        self.adam_state = None
        self.rec = {}
        self._shadow_version = None
        self.manual_version = 0      # bumped by every write this runtime knows about (Adam, load_state_dict, broadcast)
        self.bn_version = 0          # bumped by every training-mode BatchNorm launch (running statistics move)
        self._folded = {}
        self._folded_bn = {}
        self._plist = None
        # load_state_dict copies into the flat views without touching any counter this runtime could see
        net.register_load_state_dict_post_hook(lambda module, incompatible: self.mark_dirty())

    # ------------------------------------------------------------------------------------------------
    def _params(self):
        return list(self.net.parameters())

    def materialized(self):
        if self.flat is None:
            return False
        ps = self._params()
        lo, hi = self.flat.data_ptr(), self.flat.data_ptr() + self.flat.numel() * 4
        return all(lo <= p.data_ptr() < hi for p in (ps[0], ps[-1]))

    def materialize(self, device):
        """Move the parameters into one flat fp32 buffer on `device` (idempotent)."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("mpgan networks run on CUDA (B200) only: there is no CPU fallback")
        ps = self._params()
        offs, total = [], 0
        for p in ps:
            offs.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        flat = torch.zeros(total, dtype=torch.float32, device=device)
        grad = torch.zeros(total, dtype=torch.float32, device=device)
        old_grads = [p.grad for p in ps]
        with torch.no_grad():
            for p, off in zip(ps, offs):
                n = p.numel()
                if p.dim() >= 3:  # conv weight (d0, d1, *k) -> native (d0, *k, d1)
                    perm = (0,) + tuple(range(2, p.dim())) + (1,)
                    inv = (0, p.dim() - 1) + tuple(range(1, p.dim() - 1))
                    nshape = tuple(p.shape[i] for i in perm)
                    src = p.detach().to(device=device, dtype=torch.float32).permute(perm).contiguous()
                    flat[off:off + n].view(nshape).copy_(src)
                    p.data = flat[off:off + n].view(nshape).permute(inv)
                    p.grad = grad[off:off + n].view(nshape).permute(inv)
                else:
                    flat[off:off + n].view(p.shape).copy_(p.detach().to(device=device, dtype=torch.float32))
                    p.data = flat[off:off + n].view(p.shape)
                    p.grad = grad[off:off + n].view(p.shape)
            for p, g in zip(ps, old_grads):
                if g is not None:
                    p.grad.copy_(g.to(device))
        for b in self.net.buffers():
            if b.device != device:
                b.data = b.data.to(device)
        self.device, self.flat, self.grad = device, flat, grad
        self.offsets = {id(p): off for p, off in zip(ps, offs)}
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.adam_state = torch.zeros(4, dtype=torch.float32, device=device)
        self.shadow = torch.zeros(total, dtype=torch.bfloat16, device=device) if self.dtype == torch.bfloat16 else None
        self._build_records()
        self._shadow_version = None
        self._plist = None
        self._folded, self._folded_bn = {}, {}

    def _native(self, buf, p):
        off, n = self.offsets[id(p)], p.numel()
        return buf[off:off + n]

    def logical_view(self, buf, p):
        """View of a flat buffer (moments, gradients) with parameter ``p``'s logical shape (conv weights: the permuted
        view of the native [d0][taps][d1] block, like ``p`` itself)."""
        flat = self._native(buf, p)
        if p.dim() >= 3:
            perm = (0,) + tuple(range(2, p.dim())) + (1,)
            inv = (0, p.dim() - 1) + tuple(range(1, p.dim() - 1))
            return flat.view(tuple(p.shape[i] for i in perm)).permute(inv)
        return flat.view(p.shape)

    def _build_records(self):
        self.rec = {}
        for m in self.net.modules():
            if isinstance(m, _CONV_TYPES):
                r = ConvRec()
                r.spec = ops.ConvSpec.from_module(m)
                r.w32 = self._native(self.flat, m.weight)
                r.w = self._native(self.shadow, m.weight) if self.shadow is not None else r.w32
                r.bias = m.bias
                r.dw = self._native(self.grad, m.weight)
                r.db = m.bias.grad if m.bias is not None else None
                r.wt = None
                # a transposed bf16 copy is needed wherever the tcgen05 kernel runs in the Y -> X direction
                r.need_wt = self.shadow is not None and r.spec.cx % 16 == 0 and r.spec.cy % 16 == 0
                if r.need_wt:
                    r.wt = torch.empty(m.weight.numel(), dtype=torch.bfloat16, device=self.device)
                self.rec[m] = r
        # one descriptor table for all transposed copies: refreshed by a single launch
        wt = [r for r in self.rec.values() if r.wt is not None]
        self._wt_table = None
        if wt:
            rows = [[r.w.data_ptr(), r.wt.data_ptr(), r.spec.cy, r.spec.taps, r.spec.cx] for r in wt]
            self._wt_table = torch.tensor(rows, dtype=torch.int64, device=self.device)
            self._wt_max = max(r.spec.cy * r.spec.taps * r.spec.cx for r in wt)

    # ------------------------------------------------------------------------------------------------
    def ensure(self, device):
        if not self.materialized() or self.device != torch.device(device):
            self.materialize(device)
        self.refresh_shadows()

    def mark_dirty(self):
        """Call after writing parameters or BatchNorm buffers behind the runtime's back (``p.data`` edits)."""
        self.manual_version += 1
        self.bn_version += 1

    def _weights_version(self):
        """Everything that can tell that the master weights changed since the shadows were cast.  The flat buffer's
        own ``_version`` is useless here: the parameters are views re-pointed with ``p.data = ...``, so in-place
        parameter writes (``p.copy_``, ``nn.init.*``, ``p.mul_``) bump the PARAMETER's counter, not the buffer's.
        ``p.data.<op>_()`` edits bump nothing -- those need ``mark_dirty()``."""
        if self._plist is None:
            self._plist = self._params()
        return (self.manual_version, sum(p._version for p in self._plist))

    def refresh_shadows(self, force=False):
        """bf16 copies of the master weights (whole buffer in one launch) + transposed conv copies."""
        if self.shadow is None:
            return
        ver = self._weights_version()
        if not force and ver == self._shadow_version:
            return
        ops.cast(self.flat, self.shadow)
        self._refresh_transposes()
        self._shadow_version = ver

    def folded(self, conv, bn):
        """Inference: BatchNorm (running statistics) folded into the convolution -- bf16 weights in the layout the
        tcgen05 kernel of that layer reads ([cy][taps][cx], or the transposed [cx][taps][cy] for ConvTranspose) and an
        fp32 bias.  Recomputed only when the weights or the running statistics change (host-side glue, a handful
        of tiny torch kernels per layer per weight version)."""
        r = self.rec[conv]
        self._folded_bn[conv] = bn
        # bn_version: the BatchNorm kernels update the running statistics through raw pointers (no tensor counter moves)
        ver = (self._weights_version(), self.bn_version, bn.running_mean._version, bn.running_var._version)
        hit = self._folded.get(conv)
        if hit is not None and hit[0] == ver:
            return hit[1], hit[2]
        sp = r.spec
        with torch.no_grad():
            scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
            w3 = r.w32.view(sp.cy, sp.taps, sp.cx)
            b0 = conv.bias.detach().float() if conv.bias is not None else torch.zeros_like(scale)
            if sp.transposed:   # ConvTranspose: output channels are the cx axis; kernel wants [cx][taps][cy]
                wf = (w3 * scale.view(1, 1, -1)).permute(2, 1, 0).contiguous().to(torch.bfloat16)
            else:
                wf = (w3 * scale.view(-1, 1, 1)).contiguous().to(torch.bfloat16)
            bf = ((b0 - bn.running_mean.float()) * scale + bn.bias.detach().float()).contiguous()
        if hit is not None and hit[1].shape == wf.shape:
            # refresh IN PLACE: a captured CUDA graph (inference.GraphedGenerator) holds these pointers
            hit[1].copy_(wf), hit[2].copy_(bf)
            wf, bf = hit[1], hit[2]
        self._folded[conv] = (ver, wf, bf)
        return wf, bf

    def refold_all(self):
        """Recompute every cached folded (weights, bias) pair in place (same device pointers)."""
        for conv in list(self._folded.keys()):
            bn = self._folded_bn.get(conv)
            if bn is not None:
                self.folded(conv, bn)

    def _refresh_transposes(self):
        if self._wt_table is not None:
            ops.weight_transpose_batch(self._wt_table, self._wt_table.shape[0], self._wt_max)

    def zero_grad(self):
        if self.grad is not None:
            self.grad.zero_()

    def adam_step(self, lr, b1, b2, eps=1e-8):
        ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, lr, b1, b2, eps, self.adam_state, self.shadow)
        self.manual_version += 1
        if self.shadow is not None:  # shadow was written by the Adam kernel; only the transposes remain
            self._refresh_transposes()
            self._shadow_version = self._weights_version()

    def requires_grad(self):
        return any(p.requires_grad for p in self.net.parameters())


class FlatAdam(torch.optim.Optimizer):
    """``torch.optim.Adam`` over one network (GAN_final.py:298-308) as ONE fused launch on the flat buffers.

    A real ``torch.optim.Optimizer``: ``param_groups`` hold the network's own ``nn.Parameter``s (pytorch-lightning's
    ``toggle_optimizer`` flips ``requires_grad`` through them), ``state_dict()`` / ``load_state_dict()`` speak
    ``torch.optim.Adam``'s per-parameter layout (``step``, ``exp_avg``, ``exp_avg_sq`` in the logical NC[D]HW weight
    shapes), so optimizer states round-trip with the reference's checkpoints.  Hyper-parameters are read from
    ``param_groups[0]`` at every step (lr schedulers work); per-group differences are not supported."""

    def __init__(self, net, lr, betas=(0.9, 0.999), eps=1e-8):
        self.net = net
        super().__init__(list(net.parameters()), dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False))
        self.lr, self.betas, self.eps = lr, betas, eps

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        rt = self.net.runtime
        if rt.flat is None:
            raise RuntimeError("optimizer.step() before the network ran on a CUDA device")
        g = self.param_groups[0]
        if g.get("weight_decay", 0) or g.get("amsgrad", False):
            raise RuntimeError("FlatAdam implements the reference's plain Adam (no weight decay / amsgrad)")
        rt.adam_step(g["lr"], g["betas"][0], g["betas"][1], g["eps"])
        return loss

    def zero_grad(self, set_to_none=False):
        rt = self.net.runtime
        if rt.grad is not None:
            rt.zero_grad()

    # ---- torch.optim.Adam wire format -------------------------------------------------------------------------
    def state_dict(self):
        rt = self.net.runtime
        params = self.param_groups[0]["params"]
        state = {}
        if rt.flat is not None:
            step = int(rt.adam_state[0:1].view(torch.int32).item())
            if step > 0:
                for i, p in enumerate(params):
                    state[i] = {"step": torch.tensor(float(step)),
                                "exp_avg": rt.logical_view(rt.exp_avg, p).detach().clone(),
                                "exp_avg_sq": rt.logical_view(rt.exp_avg_sq, p).detach().clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(params)))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, state_dict):
        rt = self.net.runtime
        params = self.param_groups[0]["params"]
        groups = state_dict.get("param_groups", [])
        if groups:
            if len(groups[0]["params"]) != len(params):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
            for k, v in groups[0].items():
                if k != "params":
                    self.param_groups[0][k] = v
        st = state_dict.get("state", {})
        if not st:   # a fresh optimizer: zero moments, step 0
            if rt.flat is not None:
                rt.exp_avg.zero_(), rt.exp_avg_sq.zero_(), rt.adam_state.zero_()
            return
        if rt.flat is None:
            raise RuntimeError("load the optimizer state after the network has been moved to its CUDA device "
                               "(net.runtime.ensure(device))")
        steps = set()
        with torch.no_grad():
            for i, p in enumerate(params):
                e = st.get(i, st.get(str(i)))
                if e is None:
                    continue
                rt.logical_view(rt.exp_avg, p).copy_(e["exp_avg"].to(rt.device))
                rt.logical_view(rt.exp_avg_sq, p).copy_(e["exp_avg_sq"].to(rt.device))
                steps.add(int(float(e["step"])))
            if len(steps) > 1:
                raise ValueError(f"FlatAdam keeps one step counter per network; the loaded state has {sorted(steps)}")
            if steps:
                rt.adam_state[0:1].view(torch.int32).fill_(steps.pop())
