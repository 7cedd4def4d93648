"""profiles/<tag>_tensor_pipe.json from an `ncu --set full` raw page (csv): per C-ABI entry point, the time-weighted
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed of its kernels in the capture (bench.py copies it into the
per-kernel objects as `tensor_pipe_pct_ncu`; it is a committed profile figure, not measured in the bench run).
    python scripts/tensor_pipe.py profiles/r2_i_ncu_full_block_kernels_raw.csv profiles/r2_tensor_pipe.json"""
import csv
import json
import sys

API = {"conv_gemm_tc_kernel": "avdf_conv_gemm", "mlp_fused_kernel": "avdf_mlp_fused", "attention_banded_mma_kernel": "avdf_attention"}
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ik, it, ip = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum"), hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
acc, per = {}, []
for r in rows[2:]:
    name = r[ik]
    api = next((v for k, v in API.items() if k in name), None)
    if api is None:
        continue
    t, pct = float(r[it].replace(",", "")), float(r[ip].replace(",", ""))
    a = acc.setdefault(api, [0.0, 0.0])
    a[0] += t * pct; a[1] += t
    per.append({"kernel": name.split("(")[0][:90], "us": t, "tensor_pipe_pct": pct})
out = {"source": sys.argv[1] + " (ncu --set full --clock-control none, cold caches, serialised: per-launch tensor-pipe activity over the "
                             "launch's elapsed cycles, time-weighted per entry point)",
       "kernels": {k: round(v[0] / v[1], 2) for k, v in acc.items()}, "launches": per}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out["kernels"]))
