"""Cost model behind DESIGN.md section 9 item 7: SURVEY Appendix B option F (relation matrices folded into per-<type,
relation> weights: what is built) against option D (K,V [N,2d], relation matrices applied on-chip per <dst, relation>
group) for pyHGT's REAL ogbn-mag recipe — reverse relations and `self` added (ogbn-mag/preprocess_ogbn_mag.py:33-42), so
R = 9 and a source type serves up to 5 relations.  Pure arithmetic on the algorithmic bytes / flops of SURVEY 8(d) with the
throughputs measured in round 2 (bench.py lines under profiles/); no GPU needed.

    python scripts/option_d_model.py
"""
NODES = {"paper": 736_389, "author": 1_134_649, "institution": 8_740, "field": 59_965}
RAW = {"writes": ("author", "paper", 7_145_660), "cites": ("paper", "paper", 5_416_271),
       "has_topic": ("paper", "field", 7_505_078), "affiliated": ("author", "institution", 1_043_998)}
HBM_EDGE = 0.87 * 6567.7e9        # edge kernel, measured fraction of the measured copy bandwidth (C2)
TC_EFF = 0.58 * 1713.4e12         # typed GEMM incl. the split pass, measured fraction of the cuBLAS bf16 peak (C2)
SIMT_FP32 = 148 * 128 * 2 * 1.9e9
SMEM_BW = 148 * 128 * 1.9e9


def recipe():
    rel = []
    for name, (s, t, e) in RAW.items():
        rel.append((name, s, t, e))
        rel.append(("rev_" + name, t, s, e))
    for t, n in NODES.items():
        rel.append(("self", t, t, n))
    return rel


def main():
    rel = recipe()
    N = sum(NODES.values())
    E = sum(r[3] for r in rel)
    s_tau = {t: len({r[0] for r in rel if r[1] == t}) for t in NODES}          # relations a type is SOURCE of
    in_rel = {t: len({r[0] for r in rel if r[2] == t}) for t in NODES}         # relations arriving at a type
    rows_f = sum(NODES[t] * s_tau[t] for t in NODES)
    groups = sum(NODES[t] * in_rel[t] for t in NODES)                          # upper bound of <dst, relation> groups
    print("R=9 MAG recipe: N=%d E=%d  S_tau=%s  K'/V' rows (F)=%.2fM  <dst,rel> groups (D)<=%.2fM" %
          (N, E, s_tau, rows_f / 1e6, groups / 1e6))
    for d, H in ((256, 8), (512, 8)):
        dk = d // H
        edge_bytes = E * (2 * d * 4 + 5) + N * (2 * d * 4 + 4)
        t_edge = edge_bytes / HBM_EDGE
        upd = 3 * N * d * 4 / 6.9e12
        out = {}
        for opt, kv_rows in (("F", rows_f), ("D", N)):
            flops = 2 * d * d * (N + 2 * kv_rows + N) * 3                      # Q + K,V tables + a_linear, 3 bf16 products
            t_gemm = flops / TC_EFF
            tab = 2 * kv_rows * d * 4
            extra = ""
            if opt == "D":
                onchip = groups * 2 * (2 * H * dk * dk)                        # Q' = A_r Q and acc.M_r per group
                smem = groups * 2 * H * dk * dk * 4                            # both matrices read from smem per group
                extra = "  on-chip %.2f TFLOP fp32 (%.0f%% of SIMT peak at edge-kernel speed), relation matrices " \
                        "through smem %.0f GB (%.0f%% of the smem bandwidth) unless groups are batched on tensor cores" % (
                            onchip / 1e12, 100 * onchip / t_edge / SIMT_FP32, smem / 1e9, 100 * smem / t_edge / SMEM_BW)
            out[opt] = t_gemm + t_edge + upd
            print("  d=%d option %s: tables %.1f GB, GEMM %.1f ms, edge %.1f ms, update %.1f ms => %.1f ms/layer%s" %
                  (d, opt, tab / 1e9, 1e3 * t_gemm, 1e3 * t_edge, 1e3 * upd, 1e3 * out[opt], extra))
        print("  d=%d: option D best case %.0f%% faster" % (d, 100 * (1 - out["D"] / out["F"])))


if __name__ == "__main__":
    main()
