"""Measured time of every stage of one C3 step (profiles/r1_bench_c3_f16x3.json, per-kernel CUDA-event pass) against its floor:
algorithmic bytes / the measured copy bandwidth (6.545 TB/s) for the HBM-bound stages, three-pass flops / the fp16 tensor rate the
projections reach at the power-capped clock (1.47 PFLOP/s) for the dense ones.  python profiles/floor_table.py"""
import json
import os

d = json.load(open(os.path.join(os.path.dirname(__file__), "r1_bench_c3_f16x3.json")))
cfg = d["config"]
n, nc, T, D, H = cfg["n_ntgt"], cfg["n_valid_neighbours"], 3072, 1024, 8
BW, TF = 6545e9, 1.47e15
k = d["kernels"]
ms = lambda name: k[name]["ms_per_launch"] * k[name]["launches_per_step"]
U = 2 * D * D
rows = []
# rotation, layer-1 Q|K'|V' and A, layer-2 K'|V' (all nodes), layer-2 Q and A (centres), tgt side Q|K'|V' + A of three layers
g_f = 3 * (n * U * (1 + 3 + 1 + 2) + nc * U * 2 + 3 * T * U * 4)
rows.append(("projections + rotation (tcgen05, 3 passes)", sum(ms(x) for x in k if x.startswith("linear:linear[")), g_f / TF * 1e3, "tensor"))
rows.append(("LayerNorm (+ residual)", ms("layernorm"), (n + nc + 3 * T) * D * 12 / BW * 1e3, "HBM"))
rows.append(("cluster attention, all nodes", ms("hgt_cluster_attn:nn_full"), n * 4 * D * 4 / BW * 1e3, "HBM"))
rows.append(("cluster attention, centres", ms("hgt_cluster_attn:nn_centre"), (3 * nc * 2 + 2 * nc) * D * 4 / BW * 1e3, "HBM"))
ib = nc * D * 4 + 2 * T * H * D * 4 + T * D * 4
rows.append(("inter attention (3 layers)", ms("hgt_inter_fused:inter_fused"), 3 * ib / BW * 1e3, "HBM"))
# head on every token; tails on the tokens whose (uniform) target falls in them: 14.9 % / 77.6 %
lse_f = 3 * 2 * (T * 20002 * 1024 + 0.149 * T * 40000 * 256 + 0.776 * T * 207744 * 64)
rows.append(("log-prob GEMMs (LSE epilogue)", ms("linear_lse"), lse_f / TF * 1e3, "tensor / MUFU"))
att_f = 3 * 3 * 2 * (T * T * D)                      # 3 layers x 3 passes x (QK^T + PV') x causal half x 2 flops
att_b = 3 * (8 * T * T / 2 * 4 * 2)                  # S read + P written (lower triangle, 8 heads, 4 B) per layer
rows.append(("tgt causal attention (2 GEMMs + softmax, 3 layers)",
             ms("linear_batched_f16x3:attn_qk") + ms("linear_batched_f16x3:attn_pv") + ms("causal_softmax_split"),
             att_f / TF * 1e3 + att_b / BW * 1e3, "tensor + HBM"))
rows.append(("PQ gather + decode", ms("pq_gather_decode_presplit:pq_gather_decode"), n * (128 + 8 + 4 * D) / BW * 1e3, "HBM"))
rows.append(("centre-row gathers", ms("gather_rows"), 2 * nc * D * 4 * 2 / BW * 1e3, "HBM"))
print("| Stage | measured ms | floor ms | bound |\n|---|---|---|---|")
for r in rows:
    print(f"| {r[0]} | {r[1]:.2f} | {r[2]:.2f} | {r[3]} |")
print(f"| sum of the rows (whole step: {d['ms_per_step']:.2f}) | {sum(r[1] for r in rows):.2f} | {sum(r[2] for r in rows):.2f} | |")
