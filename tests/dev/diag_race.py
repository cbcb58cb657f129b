import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymotion_b200.ops import skeleton as sk
from pymotion_b200.topologies import parents_of, synth_torch
from oracle import oracle_c
par = parents_of("body22"); dev = torch.device("cuda")
n = 1_000_000
rot, gp, off = synth_torch(n, par, dev)
want_pos, want_rotm = oracle_c.fk(rot.cpu().numpy(), gp.cpu().numpy(), off.cpu().numpy(), par)
want_pos = torch.from_numpy(want_pos).to(dev).float(); want_rotm = torch.from_numpy(want_rotm).to(dev).float()
for chunk in ("4", "8"):
    os.environ["PMB_FK_CHUNK"] = chunk
    for rep in range(6):
        pos, rotm = sk.fk(rot, gp, off, par)
        torch.cuda.synchronize()
        badp = ((pos - want_pos).abs() > 1e-4).any(-1)      # [F, J]
        badr = ((rotm - want_rotm).abs() > 1e-4).flatten(-2).any(-1)
        fr = torch.nonzero(badp.any(-1) | badr.any(-1)).flatten().cpu().numpy()
        print(f"chunk={chunk} rep={rep}: bad frames {len(fr)}", end="")
        if len(fr):
            tiles = sorted(set((fr // 32).tolist()))
            print(" tiles", tiles[:10], "frames-in-tile", [(int(f) % 32) for f in fr[:40]])
            f = int(fr[0]); print("   joints bad pos", torch.nonzero(badp[f]).flatten().tolist(), "rot", torch.nonzero(badr[f]).flatten().tolist())
            print("   blocks", [t // 4 for t in tiles[:10]], "warp", [t % 4 for t in tiles[:10]])
        else:
            print()
