"""bring-up aid for conv_tc_hx clusters: which output tiles are wrong, and do they hold another tile's values?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from cfun_b200 import ops
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
N, Ci, S, Co = 1, 16, 16, 16
x = ops.to_cl(torch.randn(N, Ci, S, S, S, device="cuda"))
w = torch.randn(Co, Ci, 3, 3, 3, device="cuda") * 0.05
with torch.no_grad():
    yr = F.conv3d(x, w, None, 1, 1)
    y = ops.conv3d(x, w, None, 1, 1)
torch.cuda.synchronize()
print("DIAG dbg", ops.tc_debug_status())
bad = []
for d in range(S):
    for wb in range(2):
        tile = d * 2 + wb
        a = y[0, :, d, :, wb * 8:(wb + 1) * 8]
        b = yr[0, :, d, :, wb * 8:(wb + 1) * 8]
        e = float((a - b).abs().max() / yr.abs().max())
        nan = int(torch.isnan(a).sum())
        if not (e < 1e-4):
            # does it match some other tile of the reference?
            match = None
            for d2 in range(S):
                for wb2 in range(2):
                    b2 = yr[0, :, d2, :, wb2 * 8:(wb2 + 1) * 8]
                    if float((a - b2).abs().max()) < 1e-3: match = d2 * 2 + wb2
            bad.append((tile, "%.2e" % e, nan, match))
print("DIAG bad tiles (tile, err, nans, matches_ref_tile):", bad)
if os.environ.get("CFUN_HX_DEBUG") and int(os.environ["CFUN_HX_DEBUG"]) & 8:
    import numpy as np
    for tile in range(4):
        d, wb = tile // 2, tile % 2
        v = y[0, :8, d, 0, wb * 8].cpu().numpy()
        print("DIAG tile", tile, "rank %.0f blk %.0f" % (v[0], v[7]), "A words", [hex(int(u)) for u in v[1:3].view(np.uint32)],
              "B words", [hex(int(u)) for u in v[3:5].view(np.uint32)], "acc", v[5:7], "ref", yr[0, :2, d, 0, wb * 8].cpu().numpy())
    for tile in range(4):
        d, wb = tile // 2, tile % 2
        xb = x[0, :4, d, 0, wb * 8].to(torch.bfloat16).view(torch.int16).cpu().numpy().astype(np.uint16)
        print("DIAG tile", tile, "expected A words", hex(int(xb[0]) | (int(xb[1]) << 16)), hex(int(xb[2]) | (int(xb[3]) << 16)))
