"""Tile-granular LRU model of the L2 for the block kernel's neighbour-tile traffic (CPU only).

A tile = one prefix configuration (A = L-15 sites); one H.psi touches, per tile, the tile itself and
the partner tile of every active prefix bond (+ the prefix|mid crossing partner, half of it).  The model
replays that access stream for a given TILE ORDER against a byte-weighted LRU of capacity C and reports
the DRAM read volume, to compare traversal orders before spending GPU time on them.

    python scripts/l2_sim.py [L] [cap_MB ...]
"""
import sys
from collections import OrderedDict
from math import comb

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
caps = [float(x) for x in sys.argv[2:]] or [40, 60, 80, 100, 126]
k, B = L // 2, 15
A = L - B


def pad(js):
    if js < 0 or js > B:
        return 0
    M, T = 10, 5
    run = 0
    for jt in range(T + 1):
        jm = js - jt
        if 0 <= jm <= M:
            run += ((comb(M, jm) + 3) // 4 * 4) * comb(T, jt)
    return (run + 15) // 16 * 16 * 8          # bytes, f64


def tiles_rank_order():
    """prefix bit patterns in rank order ("1 first", site 0 most significant)."""
    out = []
    for key in range(1 << A):
        Pb = 0
        for q in range(A):
            if not (key >> (A - 1 - q)) & 1:
                Pb |= 1 << q
        js = k - bin(Pb).count("1")
        if 0 <= js <= B:
            out.append(Pb)
    return out


def neighbours(Pb):
    res = []
    for q in range(A - 1):
        if ((Pb >> q) ^ (Pb >> (q + 1))) & 1:
            res.append((Pb ^ (3 << q), 1.0))
    x = Pb ^ (1 << (A - 1))                      # crossing partner: about half of its elements are read
    if 0 <= k - bin(x).count("1") <= B:
        res.append((x, 0.5))
    return res


def simulate(order, cap_bytes):
    size = {Pb: pad(k - bin(Pb).count("1")) for Pb in order}
    lru, used, miss_bytes, hit_bytes = OrderedDict(), 0, 0.0, 0.0
    def touch(t, frac):
        nonlocal used, miss_bytes, hit_bytes
        s = size[t]
        if t in lru:
            lru.move_to_end(t)
            hit_bytes += s * frac
            return
        miss_bytes += s * frac
        lru[t] = s
        used += s
        while used > cap_bytes:
            _, s0 = lru.popitem(last=False)
            used -= s0
    for t in order:
        touch(t, 1.0)
        for n, f in neighbours(t):
            touch(n, f)
    return miss_bytes, hit_bytes


def order_by_sites(sig):
    """tile order with prefix sites listed from most to least significant in `sig` ("1 first" per site)."""
    base = tiles_rank_order()
    def keyf(Pb):
        v = 0
        for q in sig:
            v = (v << 1) | (0 if (Pb >> q) & 1 else 1)
        return v
    return sorted(base, key=keyf)


def grouped_greedy_order(e):
    """The order sd_blk_tile_order (sd_blk_host.h) builds: top e sites slow, popcount groups from full to
    empty, inside a group a greedy chain of adjacent-swap partners; the other prefix sites in rank order."""
    e = max(1, min(e, A))
    def lex(c, n):
        v = 0
        for q in range(n):
            v = (v << 1) | (0 if (c >> q) & 1 else 1)
        return v
    slow, nxt = {}, 0
    for p in range(e, -1, -1):
        cfgs = sorted([c for c in range(1 << e) if bin(c).count("1") == p], key=lambda c: lex(c, e))
        left = set(cfgs)
        cur = cfgs[0]
        while True:
            left.discard(cur)
            slow[cur] = nxt
            nxt += 1
            if not left:
                break
            for q in range(e - 2, -1, -1):
                if ((cur >> q) ^ (cur >> (q + 1))) & 1 and (cur ^ (3 << q)) in left:
                    cur ^= 3 << q
                    break
            else:
                cur = min(left, key=lambda c: lex(c, e))
    return sorted(tiles_rank_order(), key=lambda Pb: (slow[Pb & ((1 << e) - 1)], lex(Pb >> e, A - e)))


def bfs_order(e):
    """mode 2 of sd_blk_tile_order: breadth-first order of every popcount group of the top-e adjacent-swap graph."""
    from collections import deque
    e = max(1, min(e, A, 24))
    def lex(c, n):
        v = 0
        for q in range(n):
            v = (v << 1) | (0 if (c >> q) & 1 else 1)
        return v
    slow, nxt = {}, 0
    for p in range(e, -1, -1):
        cfgs = sorted([c for c in range(1 << e) if bin(c).count("1") == p], key=lambda c: lex(c, e))
        seen = set()
        for start in cfgs:
            if start in seen:
                continue
            dq = deque([start])
            seen.add(start)
            while dq:
                c = dq.popleft()
                slow[c] = nxt
                nxt += 1
                for q in range(e - 2, -1, -1):
                    if ((c >> q) ^ (c >> (q + 1))) & 1:
                        n = c ^ (3 << q)
                        if n not in seen:
                            seen.add(n)
                            dq.append(n)
    return sorted(tiles_rank_order(), key=lambda Pb: (slow[Pb & ((1 << e) - 1)], lex(Pb >> e, A - e)))


if __name__ == "__main__":
    base = tiles_rank_order()
    N = comb(L, k)
    tot = sum(pad(k - bin(p).count("1")) for p in base)
    print(f"L={L}: {len(base)} tiles, stored {tot / 1e9:.3f} GB, logical {N * 8 / 1e9:.3f} GB")
    orders = {"rank order (sites 0..A-1 most->least significant)": base}
    r = list(range(A))
    for split in (4, 5, 6, 7, 8):
        # middle sites slowest, then the top `split` sites, then the remaining low sites fastest
        mid = r[split:split + (A - split) // 2]
        low = r[split + (A - split) // 2:]
        orders[f"mid {mid[0]}..{mid[-1]} slowest, then top 0..{split - 1}, then low {low[0]}..{low[-1]}"] = order_by_sites(mid + r[:split] + low)
    orders["reverse significance (site A-1 slowest)"] = order_by_sites(r[::-1])
    orders["interleave top/low (0,A-1,1,A-2,...)"] = order_by_sites([x for pair in zip(r[:A // 2], r[::-1][:A // 2]) for x in pair] + ([r[A // 2]] if A % 2 else []))
    for e in (10, 12):
        orders[f"grouped greedy, top {e} sites slow (sd_blk_tile_order, SD_BLK_ORDER=1 SD_BLK_ORDER_E={e})"] = grouped_greedy_order(e)
    for e in (12, 14, A - 1):
        orders[f"breadth-first, top {e} sites slow (SD_BLK_ORDER=2 SD_BLK_ORDER_E={e})"] = bfs_order(e)
    print(f"{'order':75s} "+ " ".join(f"{c:>7.0f}MB" for c in caps) + "   (DRAM read GB per apply; + write %.2f GB)" % (tot / 1e9))
    for name, o in orders.items():
        row = []
        for c in caps:
            m, h = simulate(o, c * 1e6)
            row.append(m / 1e9)
        print(f"{name:75s} " + " ".join(f"{x:9.2f}" for x in row))


def simulate_bypass(order, cap_bytes, qfar):
    """as simulate(), but partner tiles of prefix bonds q < qfar are streamed past the cache
    (evict-first / no-allocate): they never hit and never displace anything."""
    size = {Pb: pad(k - bin(Pb).count("1")) for Pb in order}
    lru, used, miss = OrderedDict(), 0, 0.0
    def touch(t, frac):
        nonlocal used, miss
        s = size[t]
        if t in lru:
            lru.move_to_end(t)
            return
        miss += s * frac
        lru[t] = s
        used += s
        while used > cap_bytes:
            _, s0 = lru.popitem(last=False)
            used -= s0
    for t in order:
        touch(t, 1.0)
        for q in range(A - 1):
            if ((t >> q) ^ (t >> (q + 1))) & 1:
                n = t ^ (3 << q)
                if q < qfar:
                    if n in lru:
                        lru.move_to_end(n)
                    else:
                        miss += size[n]
                else:
                    touch(n, 1.0)
        x = t ^ (1 << (A - 1))
        if x in size:
            touch(x, 0.5)
    return miss


if __name__ == "__main__" and "--bypass" in sys.argv:
    pass
