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
        self.manual_version = 0
        self._folded = {}

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

    def _native(self, buf, p):
        off, n = self.offsets[id(p)], p.numel()
        return buf[off:off + n]

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
                r.need_wt = self.shadow is not None and r.spec.rank == 2 and r.spec.cx % 16 == 0 and r.spec.cy % 16 == 0
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
        self.manual_version += 1

    def refresh_shadows(self, force=False):
        """bf16 copies of the master weights (whole buffer in one launch) + transposed conv copies."""
        if self.shadow is None:
            return
        ver = (self.flat._version, self.manual_version)
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
        ver = (self.flat._version, self.manual_version, bn.running_mean._version, bn.running_var._version,
               bn.weight._version, bn.bias._version)
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
        self._folded[conv] = (ver, wf, bf)
        return wf, bf

    def _refresh_transposes(self):
        if self._wt_table is not None:
            ops.weight_transpose_batch(self._wt_table, self._wt_table.shape[0], self._wt_max)

    def zero_grad(self):
        if self.grad is not None:
            self.grad.zero_()

    def adam_step(self, lr, b1, b2, eps=1e-8):
        ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, lr, b1, b2, eps, self.adam_state, self.shadow)
        self.mark_dirty()
        if self.shadow is not None:  # shadow was written by the Adam kernel; only the transposes remain
            self._refresh_transposes()
            self._shadow_version = (self.flat._version, self.manual_version)

    def requires_grad(self):
        return any(p.requires_grad for p in self.net.parameters())


class FlatAdam:
    """Drop-in for torch.optim.Adam over one network: ``step`` / ``zero_grad`` run one fused launch each."""

    def __init__(self, net, lr, betas=(0.9, 0.999), eps=1e-8):
        self.net, self.lr, self.betas, self.eps = net, lr, betas, eps
        self.param_groups = [{"params": list(net.parameters()), "lr": lr, "betas": betas, "eps": eps}]

    def step(self, closure=None):
        loss = closure() if closure is not None else None
        rt = self.net.runtime
        if rt.flat is None:
            raise RuntimeError("optimizer.step() before the network ran on a CUDA device")
        g = self.param_groups[0]
        rt.adam_step(g["lr"], g["betas"][0], g["betas"][1], g["eps"])
        return loss

    def zero_grad(self, set_to_none=False):
        self.net.runtime.zero_grad()
