"""Layer schedule of the HyperSTARCOP U-Net (smp.Unet(mobilenet_v2)) over the
C-ABI kernels: forward, hand-written backward, all buffers carved from one device arena.

The graph is the reference's (model_module.py:238-251 -> segmentation_models_pytorch); the
execution is B200-first: NHWC activations, training-mode BatchNorm split into
statistics/finalize/apply so the normalise+activation runs fused in the consumer, the decoder's
nearest-x2 upsample + concat fused into the producer's store, skip features written once straight
into the decoder's concat buffers, gradients routed through channel-slice views instead of copies.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_RELU6, SC_BF16, SC_F32, call

MBV2_SETTING = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2),
                (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
DECODER_CHANNELS = (256, 128, 64, 32, 16)
BN_EPS, BN_MOMENTUM = 1e-5, 0.1


class DT:
    """NHWC device tensor view: base pointer, logical shape, channel stride."""
    __slots__ = ("ptr", "N", "H", "W", "C", "ld", "esize", "grad", "grad_pooled")

    def __init__(self, ptr, N, H, W, C, ld, esize):
        self.ptr, self.N, self.H, self.W, self.C, self.ld, self.esize = ptr, N, H, W, C, ld, esize
        self.grad = None          # DT of the same shape (may be a slice view)
        self.grad_pooled = None   # DT (N,2H,2W,C) whose 2x2 sum-pool is this tensor's gradient

    def slice(self, c0, c1):
        return DT(self.ptr + c0 * self.esize, self.N, self.H, self.W, c1 - c0, self.ld, self.esize)

    @property
    def pixels(self):
        return self.N * self.H * self.W


class Arena:
    """Chunked bump allocator over torch byte tensors.  The allocation sequence of a step is
    deterministic, so every step sees identical addresses (CUDA-graph friendly); chunks are only
    ever added, never freed."""

    CHUNK = 1 << 30

    def __init__(self, device):
        self.device = device
        self.chunks = []          # (tensor, base, size)
        self.cur = 0
        self.off = 0
        self.zeroed = 0           # chunks cleared since the last reset (zero-initialised arenas only)

    def reset(self):
        self.cur, self.off = 0, 0

    @property
    def reserved(self):
        return sum(c[2] for c in self.chunks)

    def alloc(self, nbytes, align=256):
        nbytes = max(int(nbytes), 1)
        while True:
            if self.cur >= len(self.chunks):
                size = max(self.CHUNK, (nbytes + 4095) // 4096 * 4096)
                t = torch.empty(size, dtype=torch.uint8, device=self.device)
                self.chunks.append((t, t.data_ptr(), size))
            t, base, size = self.chunks[self.cur]
            o = (self.off + align - 1) // align * align
            if o + nbytes <= size:
                self.off = o + nbytes
                return base + o
            self.cur += 1
            self.off = 0

    def zero(self, ptr, nbytes):
        for t, base, size in self.chunks:
            if base <= ptr < base + size:
                t[ptr - base:ptr - base + nbytes].zero_()
                return
        raise _lib.StarcopB200Error("pointer not in arena")


class UNetEngine:
    """Runs forward / backward of the network whose parameters are the tensors in `params`
    (name -> fp32 CUDA tensor, smp state_dict names) and buffers in `buffers`."""

    def __init__(self, params, buffers, grads, in_channels, device, compute_dtype="f32"):
        self.p, self.b, self.g = params, buffers, grads
        self.cin = in_channels
        self.device = device
        self.dtype = SC_F32 if compute_dtype == "f32" else SC_BF16
        self.esize = 4 if self.dtype == SC_F32 else 2
        self.arena = None
        self.tape = []
        self.record = False
        self.stream = 0
        self.use_tc = self.dtype == SC_BF16 and bool(_lib.load().sc_tc_supported())
        self.use_halo = os.environ.get("STARCOP_NO_HALO", "") == ""
        # weight gradients leave the critical path: they only feed the optimiser, so every wgrad kernel is issued on a
        # side stream (forked from / joined to the main stream with events, inside a CUDA graph these are just
        # edges) and overlaps the data-gradient / BatchNorm chain.  Results are unchanged: each gradient tensor has
        # one writer, and the arenas never reuse a buffer within a step.
        self.side_wgrad = os.environ.get("STARCOP_NO_SIDE_STREAM", "") == ""
        self._side = None
        self._main_obj = None
        self._arenas = {}                # "train" / "eval" -> (activation arena, zero-initialised arena)
        self.generation = 0              # bumped by every recording forward; a backward must match it
        self._tape_generation = -1
        self._packs = {}                 # (weight name, flip) -> (bf16 buffer, descriptor fields), insertion ordered
        self._pack_table = None
        self._packed_this_step = False
        self._plan_complete = False      # set once a full forward + backward has recorded every request

    # ------------------------------------------------------------------ memory
    def _select_arenas(self, which):
        if which not in self._arenas:
            a, z = Arena(self.device), Arena(self.device)
            z.CHUNK = 8 << 20
            self._arenas[which] = (a, z)
        self.arena, self.zarena = self._arenas[which]

    def begin_step(self, record=True):
        """Reset the bump allocators; the zero-initialised region (BN / reduction accumulators, wgrad tickets) is
        cleared with ONE memset per chunk instead of one per buffer.  Forwards that keep nothing for a backward
        (eval, no_grad) run in their own arenas, so they never overwrite the activations of a pending backward."""
        self._select_arenas("train" if record else "eval")
        self.arena.reset()
        self.zarena.reset()
        for t, _, _ in self.zarena.chunks:
            t.zero_()
        self.zarena.zeroed = len(self.zarena.chunks)

    def new(self, N, H, W, C, ld=None):
        ld = ld or C
        return DT(self.arena.alloc(N * H * W * ld * self.esize), N, H, W, C, ld, self.esize)

    def new_like(self, t):
        return self.new(t.N, t.H, t.W, t.C)

    def f32buf(self, n, zero=False, double=False):
        nb = n * (8 if double else 4)
        if zero:
            first = self.zarena.cur >= len(self.zarena.chunks)
            ptr = self.zarena.alloc(nb)
            if first or self.zarena.cur >= self.zarena.zeroed:
                # chunk created after begin_step's memset: clear it once now
                self.zarena.chunks[self.zarena.cur][0].zero_()
                self.zarena.zeroed = self.zarena.cur + 1
            return ptr
        return self.arena.alloc(nb)

    # ------------------------------------------------------------------ side stream for weight gradients
    def _wgrad_stream(self):
        """stream handle for a weight-gradient kernel: a side stream, made to wait for all work issued on the
        main stream so far (the kernel's inputs), or the main stream when the overlap is disabled.  With
        STARCOP_SIDE_STREAMS=n > 1 successive weight gradients alternate between n side streams (independent layers:
        each gradient tensor has one writer), so two small wgrad kernels can share the SMs the main stream leaves."""
        if not self.side_wgrad:
            return self.stream
        if self._side is None:
            # default (lowest) priority: the graphed step is captured on a high-priority stream (model_module.py), so
            # the weight gradients yield to the critical path (a high-priority side stream was measured: 9.87 vs 9.50 ms)
            n = max(1, int(os.environ.get("STARCOP_SIDE_STREAMS", "2")))          # measured: 9.24 (1) -> 9.16 ms (2), 9.17 (3)
            self._sides = [torch.cuda.Stream(device=self.device, priority=0) for _ in range(n)]
            self._side = self._sides[0]
            self._side_rr = 0
        s = self._sides[self._side_rr % len(self._sides)]
        self._side_rr += 1
        s.wait_stream(self._main_obj)
        self._side_used = True
        return s.cuda_stream

    def _join_side(self):
        if self.side_wgrad and self._side is not None and getattr(self, "_side_used", False):
            for s in self._sides:
                self._main_obj.wait_stream(s)
            self._side_used = False

    # ------------------------------------------------------------------ gradient routing
    def _grad_dst(self, x):
        """-> (DT to write, accumulate flag); allocates the gradient buffer on first use."""
        if x.grad is None:
            x.grad = self.new_like(x)
            return x.grad, 0
        return x.grad, 1

    def _alias_grad(self, x, g):
        if x.grad is None:
            x.grad = g
        else:
            call("sc_add_into", g.ptr, g.ld, 0, x.grad.ptr, x.grad.ld, 1, x.N, x.H, x.W, x.C, self.dtype, self.stream)

    # ------------------------------------------------------------------ primitive layers
    def _partials(self, C):
        """buffer for BatchNorm partial-sum rows (written, never accumulated: no zeroing needed)"""
        return self.arena.alloc(_lib.load().sc_bn_partials_bytes(C))

    def _bn_forward(self, y, bn, training, sums=None):
        """sums: (partials ptr, nrows) already produced by the conv's epilogue (tcgen05 path)."""
        C, P = y.C, y.pixels
        scale, shift = self.f32buf(C), self.f32buf(C)
        mean, invstd = self.f32buf(C), self.f32buf(C)
        if training and sums is None:
            part, n = self._partials(C), ctypes.c_int(0)
            call("sc_bn_stats", y.ptr, y.ld, part, ctypes.byref(n), P, C, self.dtype, self.stream)
            sums = (part, n.value)
        part, nrows = sums if training else (0, 0)
        call("sc_bn_finalize", part, nrows, P, C, self.p[bn + ".weight"].data_ptr(),
             self.p[bn + ".bias"].data_ptr(),
             self.b[bn + ".running_mean"].data_ptr(), self.b[bn + ".running_var"].data_ptr(),
             BN_MOMENTUM, BN_EPS, int(training), scale, shift, mean, invstd, self.stream)
        return scale, shift, mean, invstd

    def _bn_act(self, y, scale, shift, act, out=None, up2=False, residual=None):
        if out is None:
            out = self.new(y.N, y.H * (2 if up2 else 1), y.W * (2 if up2 else 1), y.C)
        call("sc_bn_act", y.ptr, y.ld, scale, shift, act, residual.ptr if residual else 0,
             residual.ld if residual else 0, out.ptr, out.ld, y.N, y.H, y.W, y.C, int(up2), self.dtype, self.stream)
        return out

    def _bn_backward(self, z, y, bn, stats, act, up2):
        """gradient wrt z (plain z.grad, or pooled view when z was stored upsampled) -> dy (new DT)."""
        scale, shift, mean, invstd = stats
        C = y.C
        if up2:
            dz, pooled = z.grad_pooled, 1
        else:
            dz, pooled = z.grad, 0
        assert dz is not None, f"no gradient reached BN {bn}"
        red, n = self._partials(C), ctypes.c_int(0)
        call("sc_bn_bwd_reduce", dz.ptr, dz.ld, pooled, y.ptr, y.ld, scale, shift, mean, invstd, act, red,
             ctypes.byref(n), y.N, y.H, y.W, C, self.dtype, self.stream)
        dy = self.new_like(y)
        call("sc_bn_bwd_apply", dz.ptr, dz.ld, pooled, y.ptr, y.ld, scale, shift, mean, invstd,
             self.p[bn + ".weight"].data_ptr(), act, red, n.value, dy.ptr, dy.ld,
             self.g[bn + ".weight"].data_ptr(), self.g[bn + ".bias"].data_ptr(),
             y.N, y.H, y.W, C, self.dtype, self.stream)
        return dy

    @staticmethod
    def _tc_dims(x, k, stride):
        """(N, H, W) handed to the tcgen05 kernels.  A 1 x 1 / stride-1 convolution is a per-pixel operation, so when
        the feature map itself does not tile into 8 x 16 patches (the 8 x 8 maps of 256-pixel tiles at 1/32 resolution)
        the same pixels are presented as N*H*W/128 images of 8 x 16 (NHWC keeps them contiguous): those layers stay on
        the tensor cores instead of falling back to the fp32-FMA kernel."""
        if k == 1 and stride == 1 and (x.W % 16 or x.H % 8) and (x.N * x.H * x.W) % 128 == 0:
            return (x.N * x.H * x.W) // 128, 8, 16
        return x.N, x.H, x.W

    def _tc_ok(self, x, cin, cout, k, stride):
        """tcgen05 path: bf16 storage, stride 1, 8-aligned channels, spatial patch 8 x 16."""
        _, H, W = self._tc_dims(x, k, stride)
        ho, wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        return (self.dtype == SC_BF16 and self.use_tc and (stride == 1 or (stride == 2 and k == 3)) and k in (1, 3)
                and cout % 8 == 0 and wo % 16 == 0 and ho % 8 == 0 and x.ld % 8 == 0)

    # ------------------------------------------------------------------ packed bf16 weights
    def _packed(self, wname, k, flip, cpad_in, cpad_out):
        """Device pointer of the tensor-core packing of weight `wname` (flip=1: the dgrad filter).  The first
        training step records every request (and packs it on the spot); from then on ALL of them are re-packed
        by ONE launch at the start of each step (`_pack_all`), and this is a table lookup."""
        key = (wname, flip)
        ent = self._packs.get(key)
        w = self.p[wname]
        cout, cin = w.shape[0], w.shape[1]
        created = ent is None
        if ent is None:
            rows, cols = (cpad_in, cpad_out) if flip else (cpad_out, cpad_in)
            buf = torch.empty(rows * k * k * cols, dtype=torch.bfloat16, device=self.device)
            ent = self._packs[key] = (buf, (w.data_ptr(), buf.data_ptr(), cout, cin, k * k, flip, cpad_in, cpad_out))
            self._pack_table = None                      # plan changed: rebuild the device table
        # a packing requested for the first time (a new input size made another layer eligible for the tensor
        # path after the plan was complete) is NOT in this step's batched re-pack: fill it now
        if created or not self._packed_this_step:
            call("sc_tc_pack_weights", w.data_ptr(), ent[0].data_ptr(), cout, cin, k, k, flip, cpad_in, cpad_out, self.stream)
        return ent[0].data_ptr()

    def _pack_all(self):
        """One launch re-packing every recorded weight (sc_tc_pack_weights_batch); False until a plan exists."""
        self._packed_this_step = False
        if not self._packs or not self._plan_complete:
            return
        self._build_pack_table()
        tab, n, total = self._pack_table[:3]
        call("sc_tc_pack_weights_batch", tab.data_ptr(), n, total, self.stream)
        self._packed_this_step = True

    def _build_pack_table(self):
        """device copy of the job table; built eagerly when the plan completes (never inside a graph capture)"""
        if self._pack_table is None:
            import numpy as np
            desc = np.zeros(len(self._packs), dtype=np.dtype([("w", "<u8"), ("out", "<u8"), ("offset", "<i8"), ("cout", "<i4"),
                                                                ("cin", "<i4"), ("kk", "<i4"), ("flip", "<i4"),
                                                                ("cin_pad", "<i4"), ("cout_pad", "<i4")]))
            off = 0
            for i, (buf, d) in enumerate(self._packs.values()):
                desc[i] = (d[0], d[1], off, d[2], d[3], d[4], d[5], d[6], d[7])
                off += buf.numel()
            host = torch.from_numpy(desc.view(np.uint8).copy()).pin_memory()      # pinned: the copy is legal even under capture
            self._pack_table = (host.to(self.device, non_blocking=True), len(self._packs), off, host)
            # a CUDA graph captured earlier replays the batched re-pack with the table it saw: tables are never
            # freed (a later, longer table only ADDS jobs -- new input sizes make more layers tensor-core eligible)
            self._pack_tables_alive = getattr(self, "_pack_tables_alive", []) + [self._pack_table]

    def _halo_ok(self, x, cin, cout, k, stride):
        """thin 3x3 layers (decoder blocks 2-4): one staged halo patch per tile + resident weights (conv_tc_halo.cu)"""
        # measured on B200 (profiles/r02_layer_bench.txt): it wins where one channel count is <= 32 and the other
        # <= 96 (decoder blocks 3.conv1 80 -> 32 and its dgrad 32 -> 80, 3.conv2, 4.conv1, 4.conv2: 1.2x ... 2.6x);
        # 64 -> 64 and wider stay on the per-tap kernel (49.8 vs 43.8 us)
        return (self.use_halo and self.dtype == SC_BF16 and self.use_tc and k == 3 and stride == 1 and x.ld % 8 == 0
                and min(cin, cout) <= 32 and max(cin, cout) <= 96 and bool(_lib.load().sc_tc_halo_supported(cin, cout)))

    def _dense_fprop(self, x, wname, k, stride, want_stats=False):
        """-> (y, sums): sums is the fp64 [2*Cout] statistics buffer when the conv epilogue produced it."""
        w = self.p[wname]
        cout, cin = w.shape[0], w.shape[1]
        assert cin == x.C, (wname, cin, x.C)
        pad = k // 2
        Ho, Wo = (x.H + 2 * pad - k) // stride + 1, (x.W + 2 * pad - k) // stride + 1
        y = self.new(x.N, Ho, Wo, cout)
        if self._halo_ok(x, cin, cout, k, stride):
            lib = _lib.load()
            cpad = lib.sc_tc_halo_cin_pad(cin)
            wb = self._packed(wname, 3, 0, cpad, cout)
            part, n = (self._partials(cout), ctypes.c_int(0)) if want_stats else (0, ctypes.c_int(0))
            call("sc_tc_conv3x3_halo", x.ptr, x.ld, wb, y.ptr, y.ld, part, ctypes.byref(n), x.N, x.H, x.W, cin, cout, 0,
                 self.stream)
            return y, ((part, n.value) if want_stats else None)
        if self._tc_ok(x, cin, cout, k, stride):
            cpad = _lib.load().sc_tc_cin_pad(cin)
            wb = self._packed(wname, k, 0, cpad, cout)
            # the statistics epilogue costs ~(Cout/16) x 300 cycles per 128-pixel tile: worth fusing only
            # when the tile's K loop is long enough to hide it, else the separate pass over y is cheaper
            want_stats = want_stats and cin * k * k >= int(os.environ.get("STARCOP_STATS_MIN_K", 256))
            part, n = (self._partials(cout), ctypes.c_int(0)) if want_stats else (0, ctypes.c_int(0))
            tn, th, tw = self._tc_dims(x, k, stride)
            call("sc_tc_conv_fprop", x.ptr, x.ld, wb, y.ptr, y.ld, part, ctypes.byref(n), tn, th, tw, cin, cout,
                 k, k, stride, 0, self.stream)
            return y, ((part, n.value) if want_stats else None)
        wp = self.f32buf(w.numel())
        call("sc_pack_weights", w.data_ptr(), wp, cout, cin, k, k, 0, self.stream)
        call("sc_conv_fprop", x.ptr, x.ld, wp, 0, y.ptr, y.ld, x.N, x.H, x.W, cin, cout, k, k, stride, pad,
             self.dtype, 0, self.stream)
        return y, None

    def _dense_backward(self, x, dy, wname, k, stride, need_dx=True):
        w = self.p[wname]
        cout, cin = w.shape[0], w.shape[1]
        pad = k // 2
        tc = self._tc_ok(x, cin, cout, k, stride) and dy.ld % 8 == 0
        lib = _lib.load()
        if tc:
            # deterministic split-K: partial tiles in a workspace, summed in split order by a second kernel
            tn, th, tw = self._tc_dims(x, k, stride)
            nb = lib.sc_tc_conv_wgrad_workspace_bytes(tn, th, tw, cin, cout, k, k, stride)
            part = self.arena.alloc(nb) if nb > 0 else 0
            call("sc_tc_conv_wgrad", x.ptr, x.ld, dy.ptr, dy.ld, self.g[wname].data_ptr(), part, tn, th, tw,
                 cin, cout, k, k, stride, self._wgrad_stream())
        else:
            nb = lib.sc_conv_wgrad_workspace_bytes(x.N, x.H, x.W, cin, cout, k, k, stride, pad)
            part = self.arena.alloc(nb) if nb > 0 else 0
            call("sc_conv_wgrad", x.ptr, x.ld, dy.ptr, dy.ld, self.g[wname].data_ptr(), part, x.N, x.H, x.W, cin, cout,
                 k, k, stride, pad, self.dtype, self._wgrad_stream())
        if need_dx:
            assert stride == 1
            dst, acc = self._grad_dst(x)
            if tc and self._halo_ok(dy, cout, cin, k, 1) and dst.ld % 8 == 0:
                lib = _lib.load()
                cpad = lib.sc_tc_halo_cin_pad(cout)
                wb = self._packed(wname, 3, 1, cin, cpad)
                call("sc_tc_conv3x3_halo", dy.ptr, dy.ld, wb, dst.ptr, dst.ld, 0, 0, dy.N, dy.H, dy.W, cout, cin, acc,
                     self.stream)
            elif tc:
                cpad = _lib.load().sc_tc_cin_pad(cout)          # dgrad conv: input channels = Cout
                wb = self._packed(wname, k, 1, cin, cpad)
                tn, th, tw = self._tc_dims(dy, k, 1)
                call("sc_tc_conv_fprop", dy.ptr, dy.ld, wb, dst.ptr, dst.ld, 0, 0, tn, th, tw, cout, cin, k, k,
                     1, acc, self.stream)
            else:
                wp = self.f32buf(w.numel())
                call("sc_pack_weights", w.data_ptr(), wp, cout, cin, k, k, 1, self.stream)
                call("sc_conv_fprop", dy.ptr, dy.ld, wp, 0, dst.ptr, dst.ld, dy.N, dy.H, dy.W, cout, cin, k, k, 1, pad,
                     self.dtype, acc, self.stream)

    # ------------------------------------------------------------------ composite blocks
    def conv_bn_act(self, x, wname, bn, k, stride, act, training, out=None, up2=False, residual=None, need_dx=True):
        y, sums = self._dense_fprop(x, wname, k, stride, want_stats=training)
        stats = self._bn_forward(y, bn, training, sums)
        z = self._bn_act(y, stats[0], stats[1], act, out=out, up2=up2, residual=residual)
        if self.record:
            def bwd():
                dy = self._bn_backward(z, y, bn, stats, act, up2)
                self._dense_backward(x, dy, wname, k, stride, need_dx)
            self.tape.append(bwd)
        return z

    def _stem_s2d_ok(self, x, cout):
        return (self.use_tc and self.use_halo and self.dtype == SC_BF16 and x.C <= 4 and x.H % 2 == 0 and x.W % 2 == 0
                and os.environ.get("STARCOP_NO_STEM_S2D", "") == "" and bool(_lib.load().sc_tc_halo_supported(16, cout))
                and bool(_lib.load().sc_tc_wgrad_halo_supported(x.N, x.H // 2, x.W // 2, 16, cout)))

    def stem_conv_bn_act(self, x, wname, bn, act, training):
        """The stride-2 stem as a space-to-depth convolution (csrc/elementwise.cu, sc_stem_s2d): a stride-1 3x3 layer
        over 16 channels on the halo kernels.  Its weight gradient is the last kernel of the backward pass and sits
        fully exposed before Adam: 166 us on the per-tap kernel (nine stride-2 boxes of 16-byte pixels per stage)."""
        w = self.p[wname]
        cout = w.shape[0]
        if not self._stem_s2d_ok(x, cout):
            return self.conv_bn_act(x, wname, bn, 3, 2, act, training, need_dx=False)
        N, H2, W2 = x.N, x.H // 2, x.W // 2
        xs = self.new(N, H2, W2, 16)
        call("sc_stem_s2d", x.ptr, x.ld, x.C, xs.ptr, N, x.H, x.W, self.stream)
        if getattr(self, "_stem_wb", None) is None or self._stem_wb.numel() != cout * 144:
            self._stem_wb = torch.empty(cout * 144, dtype=torch.bfloat16, device=self.device)
        call("sc_stem_s2d_pack_weights", w.data_ptr(), self._stem_wb.data_ptr(), cout, x.C, self.stream)
        y = self.new(N, H2, W2, cout)
        part, n = (self._partials(cout), ctypes.c_int(0)) if training else (0, ctypes.c_int(0))
        call("sc_tc_conv3x3_halo", xs.ptr, 16, self._stem_wb.data_ptr(), y.ptr, y.ld, part, ctypes.byref(n), N, H2, W2, 16, cout,
             0, self.stream)
        stats = self._bn_forward(y, bn, training, (part, n.value) if training else None)
        z = self._bn_act(y, stats[0], stats[1], act)
        if self.record:
            C = x.C

            def bwd():
                dy = self._bn_backward(z, y, bn, stats, act, False)
                lib = _lib.load()
                g16 = self.f32buf(cout * 144, zero=True)
                nb = lib.sc_tc_conv_wgrad_workspace_bytes(N, H2, W2, 16, cout, 3, 3, 1)
                ws = self.arena.alloc(nb) if nb > 0 else 0
                st = self._wgrad_stream()
                call("sc_tc_conv_wgrad", xs.ptr, 16, dy.ptr, dy.ld, g16, ws, N, H2, W2, 16, cout, 3, 3, 1, st)
                call("sc_stem_s2d_unpack_grad", g16, self.g[wname].data_ptr(), cout, C, st)
            self.tape.append(bwd)
        return z

    def inverted_residual(self, x, prefix, cin, cout, stride, t, training, out=None):
        """torchvision InvertedResidual: [1x1 expand+BN+ReLU6] -> dw3x3+BN+ReLU6 -> 1x1 project+BN (+x)."""
        hidden = cin * t
        use_res = stride == 1 and cin == cout
        i = 0
        if t != 1:
            y1, sums1 = self._dense_fprop(x, f"{prefix}.conv.0.0.weight", 1, 1, want_stats=training)
            bn1 = f"{prefix}.conv.0.1"
            st1 = self._bn_forward(y1, bn1, training, sums1)
            dw_in, dw_scale, dw_shift, dw_act = y1, st1[0], st1[1], ACT_RELU6
            i = 1
        else:
            y1, st1, bn1 = None, None, None
            dw_in, dw_scale, dw_shift, dw_act = x, 0, 0, ACT_NONE
        wdw, bn2 = f"{prefix}.conv.{i}.0.weight", f"{prefix}.conv.{i}.1"
        Ho, Wo = (x.H - 1) // stride + 1, (x.W - 1) // stride + 1
        y2 = self.new(x.N, Ho, Wo, hidden)
        # the depthwise kernel emits the BatchNorm partial sums of its outputs: no statistics pass over y2
        part2, n2 = (self._partials(hidden), ctypes.c_int(0)) if training else (0, ctypes.c_int(0))
        call("sc_dwconv_fprop", dw_in.ptr, dw_in.ld, dw_scale, dw_shift, dw_act, self.p[wdw].data_ptr(),
             y2.ptr, y2.ld, part2, ctypes.byref(n2), x.N, x.H, x.W, hidden, stride, self.dtype, self.stream)
        st2 = self._bn_forward(y2, bn2, training, (part2, n2.value) if training else None)
        z2 = self._bn_act(y2, st2[0], st2[1], ACT_RELU6)
        wpr, bn3 = f"{prefix}.conv.{i + 1}.weight", f"{prefix}.conv.{i + 2}"
        y3, sums3 = self._dense_fprop(z2, wpr, 1, 1, want_stats=training)
        st3 = self._bn_forward(y3, bn3, training, sums3)
        z3 = self._bn_act(y3, st3[0], st3[1], ACT_NONE, out=out, residual=x if use_res else None)
        if self.record:
            def bwd():
                dy3 = self._bn_backward(z3, y3, bn3, st3, ACT_NONE, False)
                self._dense_backward(z2, dy3, wpr, 1, 1)
                dy2 = self._bn_backward(z2, y2, bn2, st2, ACT_RELU6, False)
                ws = self.arena.alloc(_lib.load().sc_dwconv_wgrad_workspace_bytes(hidden))
                call("sc_dwconv_wgrad", dw_in.ptr, dw_in.ld, dw_scale, dw_shift, dw_act, dy2.ptr, dy2.ld,
                     self.g[wdw].data_ptr(), ws, x.N, x.H, x.W, hidden, stride, self.dtype, self._wgrad_stream())
                if use_res:
                    self._alias_grad(x, z3.grad)        # d(x + f(x)) -> x gets dz3 as is
                if t != 1:
                    z1 = DT(0, y1.N, y1.H, y1.W, y1.C, y1.C, self.esize)   # virtual: only its gradient exists
                    z1.grad = self.new_like(y1)
                    call("sc_dwconv_dgrad", dy2.ptr, dy2.ld, self.p[wdw].data_ptr(), z1.grad.ptr, z1.grad.ld,
                         x.N, x.H, x.W, hidden, stride, self.dtype, self.stream)
                    dy1 = self._bn_backward(z1, y1, bn1, st1, ACT_RELU6, False)
                    self._dense_backward(x, dy1, f"{prefix}.conv.0.0.weight", 1, 1)
                else:
                    # depthwise reads x directly: its data gradient goes to x
                    if x.grad is None:
                        x.grad = self.new_like(x)
                        call("sc_dwconv_dgrad", dy2.ptr, dy2.ld, self.p[wdw].data_ptr(), x.grad.ptr, x.grad.ld,
                             x.N, x.H, x.W, hidden, stride, self.dtype, self.stream)
                    else:
                        tmp = self.new_like(x)
                        call("sc_dwconv_dgrad", dy2.ptr, dy2.ld, self.p[wdw].data_ptr(), tmp.ptr, tmp.ld,
                             x.N, x.H, x.W, hidden, stride, self.dtype, self.stream)
                        self._alias_grad(x, tmp)
            self.tape.append(bwd)
        return z3

    # ------------------------------------------------------------------ whole network
    def forward(self, x_nhwc, logits_ptr, training, record=None):
        """x_nhwc: DT (N,H,W,Cin) normalised input; writes (N,H,W) fp32 logits.
        training: BatchNorm uses batch statistics and updates the running ones;
        record: keep what the backward pass needs (defaults to `training`)."""
        self.record = training if record is None else record
        if self.record and not training:
            raise _lib.StarcopB200Error("backward through eval-mode BatchNorm is not implemented: call .train() before a step that needs gradients")
        N, H, W = x_nhwc.N, x_nhwc.H, x_nhwc.W
        assert H % 32 == 0 and W % 32 == 0, "input height and width must be divisible by 32 (smp check_input_shape)"
        tape_keep = (self.tape, getattr(self, "_head_in", None))
        if self.record:
            self.generation += 1
            self._tape_generation = self.generation
        self.tape = []
        if self.use_tc:
            self._pack_all()
        E = "encoder.features"
        skip_c = (16, 24, 32, 96)                          # f1..f4
        # decoder concat buffers: [upsampled | skip], block i works at stride 16 >> i
        cat = []
        up_c = (1280,) + DECODER_CHANNELS[:-1]
        for i in range(5):
            s = 16 >> i
            cs = skip_c[3 - i] if i < 4 else 0
            cat.append(self.new(N, H // s, W // s, up_c[i] + cs))
        skip_dst = {1: cat[3].slice(up_c[3], up_c[3] + 16), 3: cat[2].slice(up_c[2], up_c[2] + 24),
                    6: cat[1].slice(up_c[1], up_c[1] + 32), 13: cat[0].slice(up_c[0], up_c[0] + 96)}

        skip_obj, up_src = {}, {}
        # ---- encoder
        x = self.stem_conv_bn_act(x_nhwc, f"{E}.0.0.weight", f"{E}.0.1", ACT_RELU6, training)
        idx, cin = 1, 32
        for t, c, n, s in MBV2_SETTING:
            for r in range(n):
                x = self.inverted_residual(x, f"{E}.{idx}", cin, c, s if r == 0 else 1, t, training,
                                           out=skip_dst.get(idx))
                if idx in skip_dst:
                    skip_obj[idx] = x
                cin = c
                idx += 1
        # features.18 (1x1 320->1280) stored upsampled x2 straight into cat[0][:, :1280]
        up_src[0] = self.conv_bn_act(x, f"{E}.18.0.weight", f"{E}.18.1", 1, 1, ACT_RELU6, training,
                                     out=cat[0].slice(0, 1280), up2=True)
        skip_of_block = {0: 13, 1: 6, 2: 3, 3: 1}
        self._tape_decoder_start = len(self.tape)          # backward runs the tape in reverse: decoder first
        # ---- decoder
        for i in range(5):
            D = f"decoder.blocks.{i}"
            if self.record:
                def route(ci=cat[i], cup=up_c[i], src=up_src[i], skip=skip_obj.get(skip_of_block.get(i))):
                    # split d(cat) into the pooled gradient of the upsampled producer and the skip's gradient
                    g = ci.grad
                    src.grad_pooled = g.slice(0, cup)
                    if skip is not None:
                        self._alias_grad(skip, g.slice(cup, ci.C))
                self.tape.append(route)
            z = self.conv_bn_act(cat[i], f"{D}.conv1.0.weight", f"{D}.conv1.1", 3, 1, ACT_RELU, training)
            if i < 4:
                up_src[i + 1] = self.conv_bn_act(z, f"{D}.conv2.0.weight", f"{D}.conv2.1", 3, 1, ACT_RELU, training,
                                                 out=cat[i + 1].slice(0, DECODER_CHANNELS[i]), up2=True)
            else:
                z = self.conv_bn_act(z, f"{D}.conv2.0.weight", f"{D}.conv2.1", 3, 1, ACT_RELU, training)
        # ---- head
        hw, hb = self.p["segmentation_head.0.weight"], self.p["segmentation_head.0.bias"]
        call("sc_head_fprop", z.ptr, z.ld, hw.data_ptr(), hb.data_ptr(), logits_ptr, N, H, W, z.C, self.dtype, self.stream)
        if self.record:
            self._head_in = z
        else:
            self.tape, self._head_in = tape_keep       # a pending backward keeps its tape
        return cat

    def backward(self, dlogits_ptr, generation=None):
        """generation: the value of `self.generation` right after the forward this backward belongs to."""
        if generation is not None and generation != self.generation:
            raise _lib.StarcopB200Error(
                "backward() of a forward whose activations were overwritten by a later recording forward: the "
                "engine keeps ONE set of activations (call backward before the next training forward)")
        if self._tape_generation != self.generation or not self.tape:
            raise _lib.StarcopB200Error("backward() called twice for one forward (retain_graph is not supported)")
        self._select_arenas("train")
        self._main_obj = torch.cuda.current_stream(self.device)
        z = self._head_in
        z.grad = self.new_like(z)
        hw = self.p["segmentation_head.0.weight"]
        # data gradient by the per-pixel kernel; weight / bias gradient by the TMA tile kernel (partial rows)
        call("sc_head_bwd", z.ptr, z.ld, hw.data_ptr(), dlogits_ptr, z.grad.ptr, z.grad.ld,
             z.N, z.H, z.W, z.C, self.dtype, self.stream)
        ws = self.arena.alloc(_lib.load().sc_head_wgrad_workspace_bytes(z.C))
        call("sc_head_wgrad_tiled", z.ptr, z.ld, dlogits_ptr, self.g["segmentation_head.0.weight"].data_ptr(),
             self.g["segmentation_head.0.bias"].data_ptr(), ws, z.N, z.H, z.W, z.C, self.dtype, self._wgrad_stream())
        hook = getattr(self, "on_decoder_grads_issued", None)
        for k in range(len(self.tape) - 1, -1, -1):
            self.tape[k]()
            if k == self._tape_decoder_start and hook is not None:
                # every head / decoder weight gradient has been issued (main + side stream): the data-parallel
                # exchange of that bucket can start while the encoder's backward runs
                if self.side_wgrad and self._side is not None:
                    for extra in self._sides[1:]:          # the hook's communication stream waits on ONE side stream
                        self._side.wait_stream(extra)
                hook(self._main_obj, self._side if self.side_wgrad else None)
        self._join_side()
        self.tape = []
        self._tape_generation = -1
        self._plan_complete = True          # every fprop and dgrad packing request of the network is now recorded
        if self._packs and self._pack_table is None:
            self._build_pack_table()
