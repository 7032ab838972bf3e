"""BASELINE.json configs[4]: sweep of rays/env x terrain size for the heightmap ray-cast (Camera.get_depths) on one B200.

Terrains: heightfield meshes in the benchmark's format (synth.make_world), T in {0.1M, 1M, 4M, 16M} triangles; patterns: the
reference's generator (heightmap_distribution.py:36-115) at other grid spacings, P in {273 .. 4398} rays/env.
For every (T, P): ms per launch for `--envs` envs (CUDA events, 3 pose sets cycled, warm-up first), rays/s, the
gather-level figure rays/s x 3602 B (K x 18 B of fp16 triangle data + 2 B out per ray at K = 200) and the
HBM-level algorithmic figure (index rows of the distinct cells + unique triangles + in/out bytes per env); a 48-env subsample is
checked bit for bit against the per-pair cross-check kernel (variant 1).
    python tools/sweep_c5.py [--envs 4096] [--terrains 0,1,2,3] > profiles/rN_sweep_c5.txt
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isaac_rover_b200 as R      # noqa: E402
from isaac_rover_b200.heightmap_distribution import build_pattern      # noqa: E402

TERRAINS = [  # (label, length m, heightfield vertices per side)
    ("0.1M", 63.28, 225),      # 0.2825 m vertex spacing (the benchmark's mesh density), 100 352 triangles
    ("1M", 200.0, 708),        # BASELINE configs[1]: 999 698 triangles
    ("4M", 400.0, 1416),       # 4 004 450 triangles, 16 M index cells
    ("16M", 400.0, 2830),      # 16 006 482 triangles on the same 16 M cells (0.1414 m spacing: a finer mesh)
]
PATTERNS = [(0.3, 0.15), (0.2, 0.1), (0.15, 0.0762), (0.15, 0.05), (0.1, 0.05), (0.065, 0.05)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--terrains", default="0,1,2,3")
    ap.add_argument("--reps", type=int, default=6)
    args = ap.parse_args()
    dev = "cuda:0"
    N = args.envs
    hbm_peak = 6547.8
    try:
        import json
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print("# C5 sweep: %d envs, K=200, res 0.1 m; HBM peak %.1f GB/s (measured copy)" % (N, hbm_peak))
    print("%-5s %9s %5s %9s %11s %12s %12s %9s %8s %s" % ("T", "triangles", "P", "ms", "Grays/s", "gatherGB/s", "algoHBM GB/s", "frac_hbm",
                                                          "fallback", "check"))
    for ti in [int(x) for x in args.terrains.split(",")]:
        label, length, nv = TERRAINS[ti]
        t0 = time.perf_counter()
        w = R.synth.make_world(length=length, nv=nv, K=200, n_stones=max(10, int(2000 * (length / 200.0) ** 2)), seed=42, build_index=None,
                               hm_res=0.025 if length <= 200.0 else 0.2)
        w.map_indices = R.build_knn_index(w.triangles, w.vertices, w.G, w.res, w.K, device=dev)
        cam = R.Camera(dev, torch.tensor([0, 0, 0.0]), assets=(w.map_indices, w.triangles, w.vertices))
        cam.map_indices = None
        w.map_indices = None
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        print("# terrain %s: %d triangles, %d x %d cells, layer %.2f GB, set-up %.1f s" %
              (label, w.triangles.shape[0], w.G, w.G, cam.layer.bytes() / 1e9, time.perf_counter() - t0), flush=True)
        sets = []
        for s in range(3):
            st = R.synth.make_env_state(w, N, seed=100 + s)
            pos, quat = st["pos"].to(dev), st["quat"].to(dev)
            sets.append((pos, R.tensor_quat_to_eul(quat)))
        for dc, df in PATTERNS:
            pts, ci, fi = build_pattern(delta_coarse=dc, delta_fine=df)
            P = pts.shape[0]
            cam.heightmap_distribution = torch.from_numpy(pts).to(dev)
            cam.num_exteroceptive = P
            cam.variant = 0
            for i in range(3):
                cam.get_depths(*sets[i % 3], want_pt=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.reps):
                cam.get_depths(*sets[i % 3], want_pt=False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            ms_alt = None
            if label == "16M":          # the degenerate mesh: also the route the library does NOT take by itself (shadow kernel forced)
                os.environ["RVB_SHADOW_FORCE"] = "1"
                for i in range(2):
                    cam.get_depths(*sets[i % 3], want_pt=False)
                torch.cuda.synchronize()
                e0.record()
                for i in range(3):
                    cam.get_depths(*sets[i % 3], want_pt=False)
                e1.record()
                torch.cuda.synchronize()
                ms_alt = e0.elapsed_time(e1) / 3
                os.environ.pop("RVB_SHADOW_FORCE")
            # distinct cells of the first 64 envs -> HBM-level algorithmic bytes per env (index rows + unique triangles + in/out)
            d, _, src = cam.get_depths(sets[0][0][:64], sets[0][1][:64], want_pt=True)
            cx = torch.round(torch.clamp((src[..., 0].float() - 0.0) / 0.1, 0, w.G - 1)).long()
            cy = torch.round(torch.clamp((src[..., 1].float() - 0.0) / 0.1, 0, w.G - 1)).long()
            cells = sum(torch.unique(cx[i] * w.G + cy[i]).numel() for i in range(64)) / 64.0
            algo_env = cells * 200 * 4 + 30000.0 * P / 1634.0 + 28 + P * 2
            rays_s = N * P / (ms * 1e-3)
            # bit-exactness of a subsample against the per-pair kernel
            cam.variant = 0
            d0, _, _ = cam.get_depths(sets[1][0][:48], sets[1][1][:48], want_pt=False)
            cam.variant = 1
            d1, _, _ = cam.get_depths(sets[1][0][:48], sets[1][1][:48], want_pt=False)
            ok = bool(torch.equal(d0.view(torch.int16), d1.view(torch.int16)))
            hits = float((d0.float() < 11.0).float().mean())
            print("%-5s %9d %5d %9.3f %11.3f %12.1f %12.1f %9.3f %8s %s" %
                  (label, w.triangles.shape[0], P, ms, rays_s / 1e9, rays_s * 3602 / 1e9, N * algo_env / (ms * 1e-3) / 1e9,
                   N * algo_env / (ms * 1e-3) / 1e9 / hbm_peak, "-", ("bit-exact vs per-pair (48 envs), %.1f%% rays hit, %.0f cells/env" % (100 * hits, cells) if ok else "MISMATCH") +
                   ("" if ms_alt is None else "; shadow kernel forced: %.3f ms" % ms_alt)),
                  flush=True)
        cam.layer.close()
        del cam, w
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
