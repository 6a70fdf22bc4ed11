"""Dev tool (GPU box): timeline of the pair-tile prefill kernel, CTA 0, from the -DHI_PAIR_TRACE build.

    python -m hydrainfer_b200.build --variant trace -DHI_PAIR_TRACE          # here (nvcc cross-compiles)
    HI_B200_LIB=hydrainfer_b200/lib/libhi_b200_trace.so python tools/pair_trace.py pre1k [--items 6]

Prints, per work item of CTA 0, when each role (K/V TMA producers, the two MMA warps, the two softmax warpgroups) passed its
hand-off points, in cycles relative to the first record, and a per-item summary of where the tensor pipe waited."""
from __future__ import annotations

import argparse
import ctypes
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from hydrainfer_b200 import _lib  # noqa: E402
from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd  # noqa: E402
from hydrainfer_b200.workloads import make_batch  # noqa: E402

ROLES = ["K-tma", "V-tma", "mma0", "mma1"] + [f"smx{w >> 2}w{w & 3}" for w in range(8)]  # lane 0 of each softmax warp
N_ROLES, CAP = 12, 4096
SMX0, SMX1 = 4, 8  # first softmax warp of tile 0 / 1
_SMX_TAGS = {1: "item", 3: "S", 4: "P", 5: "Ofull", 6: "epi-done", 7: "rescale", 8: "h0-packed", 9: "h0-bufree", 10: "h0-bar1", 11: "h0-bar2",
             12: "h1-packed", 13: "h1-bufree", 14: "h1-bar1", 15: "h1-bar2"}
TAGS = {0: {1: "item", 2: "Qempty", 3: "stage-free"}, 1: {1: "item", 3: "stage-free"},
        2: {1: "item", 2: "Qfull", 3: "ready", 4: "issued"}, 3: {1: "item", 2: "Qfull", 3: "ready", 4: "issued"},
        **{4 + w: _SMX_TAGS for w in range(8)}}
CASES = {
    "pre256": ([(256, 256)] * 32, 28, 4), "pre1k": ([(1024, 1024)] * 8, 28, 4), "pre4k": ([(4096, 4096)] * 2, 28, 4),
    "pre8k": ([(8192, 8192)], 28, 4), "cfg3p": ([(512, 512), (512, 2048), (512, 4096), (512, 8192)], 28, 4),
    "mha2k": ([(2048, 2048)] * 4, 32, 32),
}
_g = torch.Generator().manual_seed(0)
_cfg3_dec = [(1, int(L)) for L in torch.randint(256, 8193, (48,), generator=_g).tolist()]
CASES["cfg3mix"] = (_cfg3_dec + CASES["cfg3p"][0], 28, 4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("--items", type=int, default=6)
    ap.add_argument("--raw", action="store_true")
    args = ap.parse_args()
    seq_lens, hq, hkv = CASES[args.case]
    d, bs, dev = 128, 16, "cuda:0"
    batch = make_batch(seq_lens, hq, hkv, d, bs, dtype=torch.bfloat16, device=dev, gen_device=dev, seed=0)
    t = batch.n_tokens
    q3 = batch.query.view(t, hq, d)
    out = torch.empty_like(q3)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
    tt = int(_lib.lib.hi_attention_tile_tokens(hq, hkv))
    items = sorted(((L - q + min(q, (tile + 1) * tt), b, tile) for b, (q, L) in enumerate(seq_lens) for tile in range((q + tt - 1) // tt)), reverse=True)
    plan = i32([x for _, b, tile in items for x in (b, tile)]).view(-1, 2)
    work = sum(c for c, _, _ in items)
    meta = (i32(batch.q_cu_seq_lens), i32(batch.kv_cu_seq_lens), i32(batch.block_tables), i32(batch.cu_blocks_lens))
    for _ in range(3):
        mha_varlen_fwd(out, q3, batch.key_cache, batch.value_cache, *meta, None, batch.q_max, batch.kv_max, 1 / math.sqrt(d), 0, -1, 0, 0, 4, plan, tt, work)
    torch.cuda.synchronize()
    # ---- load balance of the item walk: per-CTA busy time, items and steps
    stats = (ctypes.c_ulonglong * (1024 * 4))()
    fs = _lib.lib.hi_debug_pair_cta_stats
    fs.argtypes = [ctypes.c_void_p]
    if fs(ctypes.byref(stats)) == 0:
        rows = [(stats[4 * i], stats[4 * i + 1], stats[4 * i + 2], stats[4 * i + 3]) for i in range(148) if stats[4 * i + 1] > 0]
        t_first = min(r[0] for r in rows)
        ends = sorted((r[1] - t_first) / 1e3 for r in rows)
        steps = sorted(r[3] for r in rows)
        print(f"CTAs {len(rows)}: end time us min {ends[0]:.1f} median {ends[len(ends) // 2]:.1f} max {ends[-1]:.1f}; steps per CTA min {steps[0]} median {steps[len(steps) // 2]} max {steps[-1]} "
              f"sum {sum(steps)}; items per CTA {min(r[2] for r in rows)}..{max(r[2] for r in rows)}")
        worst = sorted(rows, key=lambda r: r[1])[-5:]
        print("latest CTAs (end us, items, steps, us per step):", [(round((r[1] - t_first) / 1e3, 1), r[2], r[3], round((r[1] - r[0]) / 1e3 / max(r[3], 1), 3)) for r in worst])
        best = sorted(rows, key=lambda r: r[1])[:5]
        print("earliest CTAs (end us, items, steps, us per step):", [(round((r[1] - t_first) / 1e3, 1), r[2], r[3], round((r[1] - r[0]) / 1e3 / max(r[3], 1), 3)) for r in best])
    cap = CAP
    rec = (ctypes.c_ulonglong * (N_ROLES * cap))()
    cnt = (ctypes.c_uint * N_ROLES)()
    fn = _lib.lib.hi_debug_pair_trace
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    assert fn(ctypes.byref(rec), ctypes.byref(cnt)) == 0
    events = []  # (clock, role, tag, step, item)
    for r in range(N_ROLES):
        for i in range(cnt[r]):
            v = rec[r * cap + i]
            events.append((v & 0xffffffffff, r, (v >> 40) & 15, (v >> 44) & 4095, (v >> 56) & 255))
    if not events:
        print("no trace records (was the library built with -DHI_PAIR_TRACE?)")
        return
    t0 = min(e[0] for e in events)
    t_end = max(e[0] for e in events)
    n_items = max(e[4] for e in events if e[1] == 2)
    print(f"case {args.case}: CTA 0 walked {n_items} items in {t_end - t0} cycles; records per role {list(cnt)}")
    by = {}
    for c, r, tag, step, item in events:
        by.setdefault((r, item), []).append((c - t0, tag, step))
    print("item | steps | mma0: item->Qfull->first issue ... last issue (dur) | smx0: first S, last P, Ofull, epi-done | gap to next item's first PV issue")
    prev_last_issue = None
    for item in range(1, n_items + 1):
        m = sorted(by.get((2, item), []))
        s = sorted(by.get((SMX0, item), []))
        if not m:
            continue
        g = lambda evs, tag: [c for c, tg, _ in evs if tg == tag]
        it0, qf, rd, iss = g(m, 1), g(m, 2), g(m, 3), g(m, 4)
        sS, sP, sO, sE = g(s, 3), g(s, 4), g(s, 5), g(s, 6)
        steps = len(iss)
        line = f"{item:4d} | {steps:5d} | item {it0[0] if it0 else -1:8d} Qfull {qf[0] if qf else -1:8d} iss0 {iss[0] if iss else -1:8d} issN {iss[-1] if iss else -1:8d} ({(iss[-1] - it0[0]) if iss and it0 else 0:6d})"
        if sS:
            line += f" | S0 {sS[0]:8d} P0 {sP[0] if sP else -1:8d} PN {sP[-1] if sP else -1:8d} Ofull {sO[0] if sO else -1:8d} epi {sE[0] if sE else -1:8d}"
        if prev_last_issue is not None and iss:
            line += f" | bubble {iss[0] - prev_last_issue:6d}"
        prev_last_issue = iss[-1] if iss else prev_last_issue
        print(line)
        if item >= args.items and not args.raw:
            pass
    # skew between the four warps of a tile and between the tiles: when lane 0 of each softmax warp published P of step j
    print("P arrival per softmax warp (cycles; tile 0 warps 0-3 | tile 1 warps 0-3)")
    for item in range(1, min(n_items, args.items) + 1):
        steps = sorted({st for w in range(8) for _, tg, st in by.get((4 + w, item), []) if tg == 4})
        for st in steps:
            cols = []
            for w in range(8):
                c = [c for c, tg, s_ in by.get((4 + w, item), []) if tg == 4 and s_ == st]
                cols.append(f"{c[0]:7d}" if c else "      -")
            print(f"  item {item:2d} step {st:3d}: " + " ".join(cols[:4]) + " | " + " ".join(cols[4:]))
    # steady-state per-step cost inside items vs item boundaries
    total_steps = sum(1 for c, r, tag, step, item in events if r == 2 and tag == 4)
    print(f"mma0 steps {total_steps}; cycles per step overall {(t_end - t0) / max(total_steps, 1):.0f} (tensor-pipe floor 2 tiles x 512 = 1024)")
    if args.raw:
        for c, r, tag, step, item in sorted(events)[: 400 * args.items]:
            if item <= args.items:
                print(f"{c - t0:9d} {ROLES[r]:7s} item {item:3d} {TAGS[r].get(tag, tag):10s} {step}")


if __name__ == "__main__":
    main()
