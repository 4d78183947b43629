"""Model of the rotated bin walk of the decode kernels (csrc/decode_common.cuh): for every position on
the per-head task line, at which step do the CTAs that own that position for the different kv heads
process it? CPU only.

    python tools/decode_walk_sim.py [--batch 64] [--ctx 8192] [--heads 8] [--ctas 148]

Prints the spread (max - min step over the heads) with and without the rotation. At the C2 shape the
rotation brings 95 % of the positions to spread 0 (the rest belong to bins that straddle a head
boundary); front to back the mean spread is ~190 steps of ~0.85 us.
"""
import argparse
import collections
import statistics


def spreads(tb, heads, ctas, rotate, min_tiles=0):
    total = tb * heads
    p = max(-(-total // ctas), min_tiles)
    at = collections.defaultdict(list)
    for i in range(ctas):
        x0 = i * p
        n = min(p, total - x0)
        if n <= 0:
            continue
        u0 = 0
        if rotate:
            u0 = (p - (x0 % tb) % p) % p
            if u0 >= n:
                u0 = 0
        for t in range(n):
            x = x0 + (u0 + t) % n
            at[x % tb].append(t)
    return [max(v) - min(v) for v in at.values()], p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ctx", type=int, default=8192)
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--ctas", type=int, default=148)
    a = ap.parse_args()
    tb = a.batch * -(-a.ctx // 128)
    for rot in (False, True):
        s, p = spreads(tb, a.heads, a.ctas, rot)
        print(f"rotate={rot}: tiles/head {tb}, tiles/bin {p}: spread max {max(s)}, mean "
              f"{statistics.mean(s):.1f}, positions with spread 0: {sum(1 for v in s if v == 0) / len(s):.3f}")


if __name__ == "__main__":
    main()
