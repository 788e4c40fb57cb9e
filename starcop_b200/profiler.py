"""Live per-kernel timing of the train step with CUDA events (never under a profiler): every C-ABI call of one
eager step is bracketed by events on the launching stream and tagged with its ALGORITHMIC work -- the FLOPs of the
operator and the bytes of each operand / result counted once -- so bench.py can report, per kernel
family, achieved TFLOP/s / GB/s against the roofline min(FLOP / peak, bytes / bandwidth), name the time-dominant
kernel honestly, and aggregate the conv stack."""
import torch

from . import _lib

ES = {0: 4, 1: 2}


def _conv_out(h, k, stride):
    return h if stride == 1 else (h + 2 * (k // 2) - k) // 2 + 1


def _w_tc_fprop(a):
    N, H, W, cin, cout, k, _, stride, acc = a[7:16]
    ho, wo = _conv_out(H, k, stride), _conv_out(W, k, stride)
    return (2.0 * N * ho * wo * k * k * cin * cout,
            2.0 * (N * H * W * cin + N * ho * wo * cout * (2 if acc else 1)) + 2.0 * cout * k * k * cin)


def _w_halo(a):
    N, H, W, cin, cout, acc = a[7:13]
    return 2.0 * N * H * W * 9 * cin * cout, 2.0 * (N * H * W * cin + N * H * W * cout * (2 if acc else 1)) + 18.0 * cin * cout


def _w_tc_wgrad(a):
    N, H, W, cin, cout, k, _, stride = a[6:14]
    ho, wo = _conv_out(H, k, stride), _conv_out(W, k, stride)
    return 2.0 * N * ho * wo * k * k * cin * cout, 2.0 * (N * H * W * cin + N * ho * wo * cout) + 4.0 * cout * cin * k * k


def _w_conv_fprop(a):
    N, H, W, cin, cout, k, _, stride, pad, dtype, acc = a[6:17]
    ho, wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    es = ES[dtype]
    return 2.0 * N * ho * wo * k * k * cin * cout, es * (N * H * W * cin + N * ho * wo * cout * (2 if acc else 1)) + 4.0 * cout * cin * k * k


def _w_conv_wgrad(a):
    N, H, W, cin, cout, k, _, stride, pad, dtype = a[6:16]
    ho, wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    return 2.0 * N * ho * wo * k * k * cin * cout, ES[dtype] * (N * H * W * cin + N * ho * wo * cout) + 4.0 * cout * cin * k * k


def _w_dw_fprop(a):
    N, H, W, C, stride, dtype = a[10:16]
    ho, wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    return 18.0 * N * ho * wo * C, ES[dtype] * (N * H * W * C + N * ho * wo * C)


def _w_dw_dgrad(a):
    N, H, W, C, stride, dtype = a[5:11]
    ho, wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    return 18.0 * N * ho * wo * C, ES[dtype] * (N * H * W * C + N * ho * wo * C)


def _w_dw_wgrad(a):
    N, H, W, C, stride, dtype = a[9:15]
    ho, wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    return 18.0 * N * ho * wo * C, ES[dtype] * (N * H * W * C + N * ho * wo * C)


WORK = {
    "sc_normalize_pack": lambda a: (0.0, 4.0 * a[6] * a[7] * a[8] * a[9] + ES[a[12]] * a[6] * a[8] * a[9] * a[11]),
    "sc_tc_conv_fprop": _w_tc_fprop,
    "sc_tc_conv3x3_halo": _w_halo,
    "sc_tc_conv_wgrad": _w_tc_wgrad,
    "sc_conv_fprop": _w_conv_fprop,
    "sc_conv_wgrad": _w_conv_wgrad,
    "sc_dwconv_fprop": _w_dw_fprop,
    "sc_dwconv_dgrad": _w_dw_dgrad,
    "sc_dwconv_wgrad": _w_dw_wgrad,
    "sc_bn_stats": lambda a: (0.0, ES[a[6]] * a[4] * a[5]),
    "sc_bn_finalize": lambda a: (0.0, 16.0 * a[1] * a[3]),
    "sc_bn_act": lambda a: (0.0, ES[a[14]] * a[9] * a[10] * a[11] * a[12] * (1 + (4 if a[13] else 1) + (1 if a[5] else 0))),
    "sc_bn_bwd_reduce": lambda a: (0.0, ES[a[16]] * a[12] * a[13] * a[14] * a[15] * (1 + (4 if a[2] else 1))),
    "sc_bn_bwd_apply": lambda a: (0.0, ES[a[21]] * a[17] * a[18] * a[19] * a[20] * (2 + (4 if a[2] else 1))),
    "sc_add_into": lambda a: (0.0, ES[a[10]] * a[6] * a[7] * a[8] * a[9] * ((4 if a[2] else 1) + 1 + (1 if a[5] else 0))),
    "sc_head_fprop": lambda a: (18.0 * a[5] * a[6] * a[7] * a[8], a[5] * a[6] * a[7] * (ES[a[9]] * a[8] + 4.0)),
    "sc_head_bwd": lambda a: (18.0 * a[6] * a[7] * a[8] * a[9], a[6] * a[7] * a[8] * (ES[a[10]] * a[9] + 4.0)),
    "sc_head_wgrad_tiled": lambda a: (18.0 * a[6] * a[7] * a[8] * a[9], a[6] * a[7] * a[8] * (ES[a[10]] * a[9] + 4.0)),
    "sc_bce_fused": lambda a: (0.0, a[4] * a[5] * (12.0 + (4.0 if a[8] else 0.0))),
    "sc_adam_step_dev": lambda a: (0.0, 28.0 * a[4]),
    "sc_tc_pack_weights_batch": lambda a: (0.0, 6.0 * a[2]),
    "sc_tc_pack_weights": lambda a: (0.0, 6.0 * a[2] * a[3] * a[4] * a[5]),
    "sc_pack_weights": lambda a: (0.0, 8.0 * a[2] * a[3] * a[4] * a[5]),
}
TENSOR_KERNELS = ("sc_tc_conv_fprop", "sc_tc_conv3x3_halo", "sc_tc_conv_wgrad")
CONV_STACK = TENSOR_KERNELS + ("sc_conv_fprop", "sc_conv_wgrad", "sc_dwconv_fprop", "sc_dwconv_dgrad", "sc_dwconv_wgrad",
                               "sc_head_fprop", "sc_head_bwd", "sc_head_wgrad_tiled")


class StepProfile:
    """with StepProfile() as prof: model.train_step_fused(batch)  ->  prof.table(peak_tflops, peak_gbs)"""

    def __init__(self):
        self.records = []

    def __enter__(self):
        self._orig = _lib.call
        prof = self

        def timed_call(name, *args):
            stream = args[-1]
            ext = torch.cuda.ExternalStream(stream) if stream else torch.cuda.default_stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            prof._orig(name, *args)
            e1.record(ext)
            work = WORK.get(name)
            fl, by = work(args) if work else (0.0, 0.0)
            prof.records.append((name, e0, e1, fl, by))

        _lib.call = timed_call
        # modules that did `from ._lib import call` hold their own reference
        from . import engine
        self._eng_orig = engine.call
        engine.call = timed_call
        return self

    def __exit__(self, *exc):
        from . import engine
        _lib.call = self._orig
        engine.call = self._eng_orig
        torch.cuda.synchronize()
        self.rows = [(n, e0.elapsed_time(e1) * 1e-3, fl, by) for n, e0, e1, fl, by in self.records]
        return False

    def table(self, peak_tflops, peak_gbs):
        """per kernel family: launches, seconds, algorithmic FLOPs / bytes, roofline seconds, fraction of roofline"""
        fam = {}
        for n, t, fl, by in self.rows:
            f = fam.setdefault(n, {"launches": 0, "s": 0.0, "flop": 0.0, "bytes": 0.0, "roof_s": 0.0})
            f["launches"] += 1
            f["s"] += t
            f["flop"] += fl
            f["bytes"] += by
            f["roof_s"] += max(fl / (peak_tflops * 1e12), by / (peak_gbs * 1e9))
        for f in fam.values():
            f["frac_of_roofline"] = f["roof_s"] / f["s"] if f["s"] > 0 else None
            f["tflops"] = f["flop"] / f["s"] / 1e12 if f["s"] > 0 else None
            f["gbs"] = f["bytes"] / f["s"] / 1e9 if f["s"] > 0 else None
        return fam
