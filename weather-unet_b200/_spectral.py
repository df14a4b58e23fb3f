"""Spectral normalisation of all discriminator weights in a few multi-tensor launches
(wu_sn_forward / wu_sn_backward, csrc/wu_spectral.cu) instead of ~25 small PyTorch kernels per layer
and forward.  Semantics are those of torch.nn.utils.spectral_norm as the reference uses it
(nets.py:26-33, disc.py:21,24): the modules keep their `weight_orig`, `weight_u`, `weight_v`
parameters / buffers (state_dict keys unchanged); this path reads and updates them directly and
returns the normalised weights as autograd-connected fp32 tensors."""
import struct

import torch

from ._lib import call, on_tensor_device, query, stream


_REC = "<QQQQQQQQQQQQQQii"  # one SnTensor record of csrc/wu_spectral.cu (14 pointers, rows, cols)


def _upload(buf, dev):
    return torch.frombuffer(bytearray(buf), dtype=torch.uint8).pin_memory().to(dev, non_blocking=True)


class FusedSpectralNorm:
    """Holds the static work lists for one set of spectrally normalised modules."""

    def __init__(self, modules):
        self.modules = list(modules)
        self.shapes = [(m.weight_orig.shape[0], m.weight_orig[0].numel()) for m in self.modules]
        self.eps = float(self._hook(self.modules[0]).eps)
        self._tables = {}
        self._static = None

    @staticmethod
    def _hook(m):
        for h in m._forward_pre_hooks.values():
            if type(h).__name__ == "SpectralNorm":
                return h
        raise RuntimeError("module is not spectrally normalised")

    def supported(self):
        """fp32 CUDA weights, dense and row-major, one power iteration (the reference's setting)."""
        for m in self.modules:
            w = m.weight_orig
            if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()):
                return False
            h = self._hook(m)
            if h.n_power_iterations != 1 or h.dim != 0:
                return False
        return True

    def _static_chunks(self, dev):
        if self._static is not None:
            return self._static
        cols_per, rows_per = query("wu_sn_wtu_cols"), query("wu_sn_wv_rows")
        half = query("wu_sn_parts") // 2
        wtu, wv, elem, dot, bwd = bytearray(), bytearray(), bytearray(), bytearray(), bytearray()
        n = [0, 0, 0, 0, 0]
        self._dot_ranges, self._bwd_ranges = [], []
        for ti, (rows, cols) in enumerate(self.shapes):
            if (cols + cols_per - 1) // cols_per > half or (rows + rows_per - 1) // rows_per > half:
                raise RuntimeError(f"FusedSpectralNorm: weight {rows}x{cols} too large")
            for i, c0 in enumerate(range(0, cols, cols_per)):
                wtu += struct.pack("<iiii", ti, c0, min(cols_per, cols - c0), i)
                n[0] += 1
            for i, r0 in enumerate(range(0, rows, rows_per)):
                wv += struct.pack("<iiii", ti, r0, min(rows_per, rows - r0), i)
                n[1] += 1
            units = (rows * cols + 1023) // 1024
            for i, e0 in enumerate(range(0, units, 8)):  # 8192 elements per block
                elem += struct.pack("<iiii", ti, e0, min(8, units - e0), i)
                n[2] += 1
            per = max(8, (units + half - 1) // half)
            nd = 0
            d0 = n[3]
            for i, e0 in enumerate(range(0, units, per)):
                dot += struct.pack("<iiii", ti, e0, min(per, units - e0), i)
                n[3] += 1
                nd += 1
            self._dot_ranges.append((d0, n[3]))
            b0 = n[4]
            for e0 in range(0, units, 8):
                bwd += struct.pack("<iiii", ti, e0, min(8, units - e0), nd)
                n[4] += 1
            self._bwd_ranges.append((b0, n[4]))
        self._static = dict(wtu=_upload(wtu, dev), wv=_upload(wv, dev), elem=_upload(elem, dev),
                            dot=_upload(dot, dev), bwd=_upload(bwd, dev), n=n)
        return self._static

    def _table(self, key, records, dev):
        """Device copy of a pointer table; re-uploaded only when a pointer changed (the caching
        allocator hands back the same blocks iteration after iteration)."""
        t = self._tables.get(key)
        if t is None:
            if len(self._tables) >= 16:
                self._tables.clear()
            t = self._tables[key] = _upload(b"".join(records), dev)
        return t

    def packs(self, i):
        """True when weight i is a 3x3 convolution the tcgen05 kernels take (packed bf16 output)."""
        w = self.modules[i].weight_orig
        return w.dim() == 4 and tuple(w.shape[2:]) == (3, 3) and w.shape[0] % 64 == 0 and w.shape[1] % 64 == 0

    def __call__(self, training):
        """-> (ws, packed): ws[i] = W / sigma (fp32, weight_orig's shape; for the packed 3x3 convolution
        weights only an autograd handle whose values are NOT written), packed[i] = (w_fprop, w_dgrad)
        bf16 operand layouts or None."""
        n = len(self.modules)
        outs = _SNAll.apply(self, bool(training), *[m.weight_orig for m in self.modules])
        ws, extra = list(outs[:n]), list(outs[n:])
        packed, k = [], 0
        for i in range(n):
            if self.packs(i):
                packed.append((extra[k], extra[k + 1]))
                k += 2
            else:
                packed.append(None)
        return ws, packed


class _SNAll(torch.autograd.Function):

    @staticmethod
    @on_tensor_device
    def forward(ctx, sn, training, *ws):
        dev = ws[0].device
        st = sn._static_chunks(dev)
        rows_tot = sum(r for r, _ in sn.shapes)
        cols_tot = sum(c for _, c in sn.shapes)
        nparts = query("wu_sn_parts")
        nt = len(ws)
        # one scratch / snapshot buffer: [sigma | u_snap | v_snap | t | s | part]
        buf = torch.empty(nt + 2 * rows_tot + 2 * cols_tot + nt * nparts, dtype=torch.float32, device=dev)
        sig_o, us_o, vs_o = 0, nt, nt + rows_tot
        t_o, s_o = vs_o + cols_tot, vs_o + 2 * cols_tot
        p_o = s_o + rows_tot
        outs = [torch.empty_like(w) for w in ws]
        extra = []
        recs, ro, co = [], 0, 0
        base = buf.data_ptr()
        for i, (m, w, o) in enumerate(zip(sn.modules, ws, outs)):
            rows, cols = sn.shapes[i]
            wf = wd = 0
            if sn.packs(i):
                cin = w.shape[1]
                tf = torch.empty((rows, 9 * cin), dtype=torch.bfloat16, device=dev)
                td = torch.empty((cin, 9 * rows), dtype=torch.bfloat16, device=dev)
                extra += [tf, td]
                wf, wd = tf.data_ptr(), td.data_ptr()
            recs.append(struct.pack(
                _REC, w.data_ptr(), m.weight_u.data_ptr(), m.weight_v.data_ptr(),
                base + 4 * (us_o + ro), base + 4 * (vs_o + co), base + 4 * (sig_o + i),
                base + 4 * (t_o + co), base + 4 * (s_o + ro), base + 4 * (p_o + i * nparts),
                o.data_ptr(), 0, 0, wf, wd, rows, cols))
            ro += rows
            co += cols
        key = ("fwd", training) + tuple(r for r in recs)
        table = sn._table(key, recs, dev)
        n = st["n"]
        with torch.no_grad():
            call("wu_sn_forward", table.data_ptr(), nt, st["wtu"].data_ptr(), n[0], st["wv"].data_ptr(),
                 n[1], st["elem"].data_ptr(), n[2], int(training), sn.eps, stream())
        ctx.sn, ctx.buf, ctx.recs = sn, buf, recs
        ctx.n = nt
        ctx.save_for_backward(*ws)
        ctx.mark_non_differentiable(*extra)
        return tuple(outs) + tuple(extra)

    @staticmethod
    @on_tensor_device
    def backward(ctx, *grads):
        grads = grads[:ctx.n]  # the packed bf16 copies are not differentiable
        sn, ws = ctx.sn, ctx.saved_tensors
        dev = ws[0].device
        st = sn._static_chunks(dev)
        live = [i for i, g in enumerate(grads) if g is not None and ctx.needs_input_grad[2 + i]]
        if not live:
            return (None, None) + tuple(None for _ in ws)
        recs, keep, dws = [], [], [None] * len(ws)
        for i, rec in enumerate(ctx.recs):
            f = list(struct.unpack(_REC, rec))
            if i in live:
                g = grads[i].contiguous().float()
                keep.append(g)
                dws[i] = torch.empty_like(ws[i])
                f[10], f[11] = g.data_ptr(), dws[i].data_ptr()
            recs.append(struct.pack(_REC, *f))
        table = sn._table(("bwd",) + tuple(recs), recs, dev)
        if len(live) == len(ws):
            dot, nd, bw, nb = st["dot"], st["n"][3], st["bwd"], st["n"][4]
            call("wu_sn_backward", table.data_ptr(), dot.data_ptr(), nd, bw.data_ptr(), nb, stream())
        else:  # some weights without gradient: launch per live tensor over its slice of the work lists
            for i in live:
                d0, d1 = sn._dot_ranges[i]
                b0, b1 = sn._bwd_ranges[i]
                call("wu_sn_backward", table.data_ptr(), st["dot"].data_ptr() + 16 * d0, d1 - d0,
                     st["bwd"].data_ptr() + 16 * b0, b1 - b0, stream())
        return (None, None) + tuple(dws)
