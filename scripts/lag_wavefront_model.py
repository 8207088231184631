"""L1-wavefront model of the lag kernels (no GPU needed): predicted time = L1 wavefronts per SM / (1 per cycle).

A warp-level load costs one wavefront per 128-byte line it touches (B300_MICROARCH.md, "L1tex wavefront queue").
The model reproduces the measured time of the default kernel at C4 (33 ms) and registers predictions for the
experiment switches BEFORE they are measured (scripts/round2_first_call.sh):

    python scripts/lag_wavefront_model.py

Inputs per configuration: N cells, ld floats per row, mean degree, the union fraction of R consecutive
neighbour lists in spatial order (simulated on uniform points: scripts note in DESIGN.md §8) and the SM clock.
"""
import math

SMS, CLOCK_GHZ, Q = 148, 1.75, 8          # Q lanes x float4 = one 128-byte piece of a row per gather


def lines_per_piece(ld, aligned):
    """Mean number of 128-byte lines a 128-byte piece of a row touches: rows start at j*4*ld bytes."""
    if aligned or (4 * ld) % 128 == 0:
        return 1.0
    period = 128 // math.gcd(4 * ld, 128)
    return 1.0 + (period - 1) / period      # only every `period`-th row starts on a line


def span_lines(n_addr, stride_bytes):
    """Lines touched by n_addr 4-byte loads `stride_bytes` apart (mean over alignments)."""
    span = (n_addr - 1) * stride_bytes + 4
    return min(float(n_addr), span / 128.0 + 1.0)


def default_kernel(n, ld, deg, aligned):
    al = lines_per_piece(ld, aligned)
    rows_per_load = 32 // Q                                   # 4 rows per warp-level gather
    per_item = deg * (rows_per_load * al + span_lines(rows_per_load, 4 * deg))   # gathers + index loads
    per_item += 2 * rows_per_load * al + 2                    # own z row, lag store, indptr
    issue = (deg * (4 + 6) + 40) / 4.0                        # 4 FFMA + ~6 other per neighbour per warp
    items = (n / rows_per_load) * math.ceil(ld / (4 * Q)) / SMS
    return max(per_item, issue) * items


def grouped_kernel(n, ld, deg, aligned, R, union_fraction):
    al = lines_per_piece(ld, aligned)
    groups_per_load = 32 // Q
    U = union_fraction * R * deg                              # union length per group
    # gathers + word loads (one 16-byte load per four union words, one line per group)
    per_item = U * (groups_per_load * al + span_lines(groups_per_load, 4 * R * deg) / 4.0)
    per_item += 2 * groups_per_load * R * al + groups_per_load + 2
    # issue-slot floor: per union word 4R predicated FADDs + ~8 address / mask instructions per warp, ~40 per
    # row in the epilogue; four schedulers per SM issue one warp instruction per cycle each
    issue = (U * (4 * R + 8) + 40 * R) / 4.0
    items = (n / (groups_per_load * R)) * math.ceil(ld / (4 * Q)) / SMS
    return max(per_item, issue) * items


def ms(wavefronts):
    return wavefronts / (CLOCK_GHZ * 1e9) * 1e3


CONFIGS = {
    # name: (N, G, degree, {R: union fraction}) -- union fractions simulated on uniform points in the Z-curve order
    "C4 radius deg 20, 5M x 1000": (5_000_000, 1000, 20.0, {2: 0.726, 4: 0.493, 8: 0.336}),
    "C2 kNN k=15, 500k x 400": (500_000, 400, 15.0, {2: 0.756, 4: 0.533, 8: 0.374}),
    "C3 kNN k=6, 200k x 1000": (200_000, 1000, 6.0, {2: 0.852, 4: 0.677, 8: 0.524}),
}

if __name__ == "__main__":
    for name, (n, g, deg, uf) in CONFIGS.items():
        print(name)
        for aligned in (False, True):
            ld = (g + 31) // 32 * 32 if aligned else (g + 7) // 8 * 8
            tag = "aligned (SC_ROW_ALIGN=32)" if aligned else "packed"
            row = [f"default {ms(default_kernel(n, ld, deg, aligned)):6.2f} ms"]
            for R, f in uf.items():
                row.append(f"R={R} {ms(grouped_kernel(n, ld, deg, aligned, R, f)):6.2f} ms")
            hbm_ms = (8.0 * n * ld + 4.0 * n * deg + 4.0 * n) / 6532.2e9 * 1e3   # Z in, lag out, CSR: the HBM floor
            print(f"  {tag:26s} ld={ld:5d}  " + "   ".join(row) + f"   (HBM floor {hbm_ms:.2f} ms)")
