"""Group the kernel table printed by tools/profile_step.py into categories (developer tool)."""
import collections
import re
import sys

rows = []
for line in open(sys.argv[1]):
    parts = re.split(r"\s{2,}", line.strip())
    if len(parts) < 11 or not parts[-1].isdigit():
        continue
    m = re.match(r"([\d.]+)(us|ms|s)$", parts[6])
    if not m:
        continue
    rows.append((parts[0], float(m.group(1)) * {"us": 1, "ms": 1e3, "s": 1e6}[m.group(2)], int(parts[-1])))
geo = ("k_fps", "k_knn", "k_count", "k_scan", "k_fill", "k_bbox", "k_occ", "k_grid", "k_pt_rel", "k_td_rel", "k_cell",
       "k_trial", "k_scene", "k_reset", "k_flag", "k_regather", "k_hist", "k_rank")
cat = collections.OrderedDict()


def add(c, u, n):
    a = cat.setdefault(c, [0, 0])
    a[0] += u
    a[1] += n


for name, u, n in rows:
    nm = name.replace("void ", "")
    if any(nm.startswith(g) for g in geo): add("geometry (FPS, KNN, grids, rel)", u, n)
    elif nm.startswith("k_tc_wgrad"): add("tc_wgrad (linear + pt dW3)", u, n)
    elif nm.startswith("k_tc_gemm<32, false, 1>"): add("tc_gemm dy2 (pt bwd)", u, n)
    elif nm.startswith("k_tc_gemm"): add("tc_gemm linear fwd/dgrad", u, n)
    elif nm.startswith("k_skinny"): add("simt skinny gemm", u, n)
    elif nm.startswith("k_pt_bwd_main"): add("pt bwd main", u, n)
    elif nm.startswith("k_pt_bwd_dw3"): add("pt bwd dw3 (simt)", u, n)
    elif nm.startswith("k_pt_bwd_softmax"): add("pt bwd softmax", u, n)
    elif nm.startswith("k_pt_bwd_da"): add("pt bwd da", u, n)
    elif nm.startswith("k_pt_softmax"): add("pt softmax", u, n)
    elif nm.startswith("k_pt_w2"): add("pt w2", u, n)
    elif nm.startswith("k_pt_w0"): add("pt w0 stats", u, n)
    elif nm.startswith("k_pt_aggregate"): add("pt aggregate", u, n)
    elif nm.startswith("k_pt_") or nm.startswith("k_bn_"): add("pt small", u, n)
    elif nm.startswith("k_td"): add("transition down", u, n)
    elif nm.startswith("k_cbl"): add("cbl", u, n)
    elif "batch_norm" in nm: add("torch batch_norm", u, n)
    elif "gemm" in nm or "cutlass" in nm or "splitK" in nm or "gemv" in nm: add("torch/cublas gemm", u, n)
    elif "multi_tensor" in nm: add("sgd", u, n)
    elif nm.startswith("k_"): add("other ours: " + nm[:28], u, n)
    else: add("torch elementwise/fill/reduce/index/other", u, n)
tot = sum(v[0] for v in cat.values())
for c, (u, n) in sorted(cat.items(), key=lambda kv: -kv[1][0]):
    print(f"{c:45s} {u / 1e3:8.2f} ms {n:5d} launches")
print(f"{'total':45s} {tot / 1e3:8.2f} ms")
