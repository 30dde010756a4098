"""Wave-quantisation model of the forward conv GEMMs: for every 3x3 conv of the
step at BASELINE configs[1] it mirrors the launcher's tile bookkeeping (csrc/igemm.cu: BN =
min(Cout, 256), 128-row tiles, CTA pairs per pick_cg, persistent grid of `units` CTAs / pairs) and
reports how full the last round of tiles is. Pure arithmetic, no GPU.
    python scripts_dev/wave_model.py [batch] [size]"""
import math
import sys

SMS = 148


def layers(n, size):
    c = [64, 128, 256, 512, 1024]
    out = []
    h = size
    for i in range(5):
        cin = 1 if i == 0 else c[i - 1]
        out.append((f"enc{i}.a", n, h, cin, c[i]))
        out.append((f"enc{i}.b", n, h - 2, c[i], c[i]))
        h -= 4
        if i < 4:
            h //= 2
    for j in range(4):
        cp = c[4 - j]
        h *= 2
        out.append((f"up{j + 1}.a", n, h, cp, cp // 2))
        out.append((f"up{j + 1}.b", n, h - 2, cp // 2, cp // 2))
        h -= 4
    return out


def model(name, n, hin, cin, cout):
    ho = hin - 2
    m = n * ho * ho
    bn = min(cout, 256)
    n_tiles = cout // bn
    # row-run (one tile = 128 pixels of ONE output row, igemm_rr.cuh) when such tiles are >= 80 % valid
    rowrun = ho * 100 >= math.ceil(ho / 128) * 128 * 80      # rowrun_eligible() in csrc/igemm.cu
    m_tiles = n * ho * math.ceil(ho / 128) if rowrun else math.ceil(m / 128)
    kblocks = 9 * max(cin // 64, 1)
    cg = 2 if (bn >= 128 and kblocks >= 16 and m_tiles * n_tiles >= 2 * SMS) else 1
    units = max((SMS // cg // n_tiles) * n_tiles, n_tiles)
    tiles = math.ceil(m_tiles / cg) * n_tiles
    units = min(units, tiles)
    rounds = math.ceil(tiles / units)
    fill = tiles / (rounds * units)                       # share of CTA-rounds that carry a tile
    valid = m / (m_tiles * 128)                            # share of tile rows that are real pixels
    gflop = 2.0 * m * cout * 9 * cin / 1e9
    return dict(name=name, M=m, K=9 * cin, N=cout, path="row-run" if rowrun else "k-major", BN=bn, CG=cg,
                tiles=tiles, units=units, rounds=rounds, fill=fill, valid=valid, gflop=gflop)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    rows = [model(*l) for l in layers(n, size) if l[3] > 1]
    tot = sum(r["gflop"] for r in rows)
    lost = 0.0
    print(f"{'layer':8s} {'M':>9s} {'N':>5s} {'K':>6s} {'path':8s} {'BN':>4s} {'CG':>3s} {'tiles':>7s} "
          f"{'units':>6s} {'rounds':>7s} {'fill':>6s} {'valid':>6s} {'GFLOP':>8s}")
    for r in rows:
        eff = r["fill"] * r["valid"]
        lost += r["gflop"] * (1.0 / eff - 1.0)
        print(f"{r['name']:8s} {r['M']:9d} {r['N']:5d} {r['K']:6d} {r['path']:8s} {r['BN']:4d} {r['CG']:3d} "
              f"{r['tiles']:7d} {r['units']:6d} {r['rounds']:7d} {r['fill']:6.3f} {r['valid']:6.3f} "
              f"{r['gflop']:8.1f}")
    print(f"forward GEMM work {tot:.0f} GFLOP; tile slots executed beyond it (last-round tails + padded "
          f"tile rows): {lost:.0f} GFLOP-equivalents = {100 * lost / tot:.1f} % of the forward GEMM time "
          f"at equal tile speed")


if __name__ == "__main__":
    main()
