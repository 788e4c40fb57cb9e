"""One launch each of the kernels profiled for profiles/: the dominant tcgen05 conv (decoder.blocks.0.conv1 fprop),
the thin-layer halo conv (decoder.blocks.4.conv2 shape), a depthwise fprop / wgrad, the resident mag1c kernel and the
cluster ratio kernel.  Target of
    ncu --set full --clock-control none --import-source on -k regex:<name> -c <n> -o gpurun_out/<x> python scripts/profile_kernels.py <what>
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from starcop_b200 import _lib, features, mag1c, synthetic
from starcop_b200._lib import call, ACT_RELU6, SC_BF16
what = sys.argv[1] if len(sys.argv) > 1 else "all"
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
B, S = 16, 512
n = ctypes.c_int(0)
if what in ("all", "dominant"):
    from starcop_b200.model_setup import get_model
    from starcop_b200.settings import default_settings
    torch.manual_seed(0)
    m = get_model(default_settings(compute_dtype="bf16"), None).cuda()
    print(m.network.bench_dominant_kernel(B, S, iters=3))
if what in ("all", "halo"):
    cin, cout = int(os.environ.get("HALO_CIN", 16)), int(os.environ.get("HALO_COUT", 16))
    S = int(os.environ.get("HALO_S", S))
    x = torch.randn(B, S, S, cin, device="cuda").bfloat16(); y = torch.empty(B, S, S, cout, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    hp = lib.sc_tc_halo_cin_pad(cin)
    wb = torch.empty(cout * 9 * hp, dtype=torch.bfloat16, device="cuda")
    call("sc_tc_pack_weights", w.data_ptr(), wb.data_ptr(), cout, cin, 3, 3, 0, hp, cout, st)
    part = torch.empty(lib.sc_bn_partials_bytes(cout) // 8, dtype=torch.float64, device="cuda")
    for _ in range(3):
        call("sc_tc_conv3x3_halo", x.data_ptr(), cin, wb.data_ptr(), y.data_ptr(), cout, part.data_ptr(), ctypes.byref(n), B, S, S, cin, cout, 0, st)
if what in ("all", "dw"):
    C, H = 96, 256
    x = torch.randn(B, H, H, C, device="cuda").bfloat16(); y = torch.empty(B, H // 2, H // 2, C, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(C, 1, 3, 3, device="cuda") / 3; dw = torch.zeros_like(w)
    sc, sh = torch.rand(C, device="cuda") + .5, torch.randn(C, device="cuda")
    part = torch.empty(lib.sc_bn_partials_bytes(C) // 8, dtype=torch.float64, device="cuda")
    ws = torch.empty(lib.sc_dwconv_wgrad_workspace_bytes(C) // 4, device="cuda")
    for _ in range(3):
        call("sc_dwconv_fprop", x.data_ptr(), C, sc.data_ptr(), sh.data_ptr(), ACT_RELU6, w.data_ptr(), y.data_ptr(), C, part.data_ptr(), ctypes.byref(n), B, H, H, C, 2, SC_BF16, st)
        call("sc_dwconv_wgrad", x.data_ptr(), C, sc.data_ptr(), sh.data_ptr(), ACT_RELU6, y.data_ptr(), C, dw.data_ptr(), ws.data_ptr(), B, H, H, C, 2, SC_BF16, st)
if what in ("all", "mag1c"):
    t73 = synthetic.synthetic_template(73)
    cube, _, _ = synthetic.aviris_cube(2, size=512, bands=125, seed=1, template=t73)
    c = torch.from_numpy(cube).cuda()
    for it in (0, 30, 30):
        mag1c.mag1c_tiles(c, t73, slice(52, 125), num_iter=it)
if what in ("all", "ratio"):
    bg = torch.rand(8, 512, 512, device="cuda") + 0.5; sig = bg * 0.8 + 0.01 * torch.randn_like(bg)
    for _ in range(3):
        features.ratio_2c_match_c_from_sums_outlier(bg, sig)
torch.cuda.synchronize()
