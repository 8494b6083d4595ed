"""profiles/ncu_traffic.json from `ncu --set full` reports: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum)
per launch of every captured kernel, next to its ALGORITHMIC bytes at the capture's shape, and the
kernel's share of its class's launches in the solve, so that bench.py can report a class-level `roofline.traffic`
that is comparable with its live `algorithmic_per_launch`.

    python scripts/ncu_traffic.py gpurun_out/ncu_<tag>_bpcg_iteration.ncu-rep gpurun_out/ncu_<tag>_dense.ncu-rep
"""
import csv, json, re, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
N, K, P, NNZ, NDP = 8_000_000, 40, 480, 119_042_398, 16      # headline shapes: P1-FEM m = 200, block of 40, p = 480


def algorithmic(kernel: str):
    """(class, bytes, flops, launches per CG iteration or per call) for the kernels of this build"""
    if "bpcg_update_px" in kernel:
        return "bpcg_fused", 40.0 * N * K, None
    if "bpcg_update_r" in kernel:
        return "bpcg_fused", 24.0 * N * K, None
    if "spmm_lat" in kernel or "spmm_dia" in kernel:
        return "spmm", 12.0 * NNZ + 4.0 * (N + 1) + 16.0 * N * K, 2.0 * NNZ * K
    if "gram" in kernel:
        return "gram", 8.0 * N * (P + K) + 8.0 * P * K, 2.0 * N * P * K
    if "lincomb" in kernel:
        return "lincomb", 8.0 * N * (P + K) + 8.0 * P * K, 2.0 * N * P * K
    return None, None, None


def main():
    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full captures of the SHIPPED kernels "
                       "(scripts/ncu_r2.sh; P1-FEM pencil n = 8.0 M, k = 40, p = 480), with the algorithmic bytes of the same launch; "
                       "class entries are launch-weighted (BlockPCG: one update_px + one update_r per CG iteration)",
           "workload": {"m": 200, "k": K}, "kernels": {}}
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, data = rows[0], rows[2:]
        ki, ri, wi, ti = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        units = rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
        for d in data:
            name = re.sub(r"\(.*", "", d[ki]).replace("void ", "")
            cls, byt, fl = algorithmic(name)
            if cls is None:
                continue
            tr = float(d[ri]) * scale[units[ri]] + float(d[wi]) * scale[units[wi]]
            e = out["kernels"].setdefault(name, {"class": cls, "traffic_bytes_per_launch": [], "ms": [],
                                                 "algorithmic_bytes_per_launch": byt, "source": f"profiles/{Path(rep).stem}.txt"})
            e["traffic_bytes_per_launch"].append(tr); e["ms"].append(float(d[ti]) * tscale[units[ti]])
    for e in out["kernels"].values():
        e["launches_captured"] = len(e["ms"])
        e["traffic_bytes_per_launch"] = sum(e["traffic_bytes_per_launch"]) / len(e["ms"])
        e["ms"] = sum(e["ms"]) / len(e["ms"])
        e["traffic_over_algorithmic"] = round(e["traffic_bytes_per_launch"] / e["algorithmic_bytes_per_launch"], 4)
    # class level: BlockPCG streams = one update_px and one update_r per CG iteration
    px = [e for k, e in out["kernels"].items() if "bpcg_update_px" in k]
    rr = [e for k, e in out["kernels"].items() if "bpcg_update_r" in k]
    if px and rr:
        out["bpcg_fused"] = {"kernel": "bpcg_update_px_kernel + bpcg_update_r_kernel (one of each per CG iteration)",
                             "traffic_bytes_per_launch": 0.5 * (px[0]["traffic_bytes_per_launch"] + rr[0]["traffic_bytes_per_launch"]),
                             "algorithmic_bytes_per_launch": 0.5 * (px[0]["algorithmic_bytes_per_launch"] + rr[0]["algorithmic_bytes_per_launch"]),
                             "source": px[0]["source"]}
    for cls in ("spmm", "gram", "lincomb"):
        es = [(k, e) for k, e in out["kernels"].items() if e["class"] == cls]
        if es:
            k, e = max(es, key=lambda kv: kv[1]["launches_captured"])
            out[cls] = {"kernel": k, "traffic_bytes_per_launch": e["traffic_bytes_per_launch"],
                        "algorithmic_bytes_per_launch": e["algorithmic_bytes_per_launch"], "source": e["source"]}
    (ROOT / "profiles" / "ncu_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
    for k, e in out["kernels"].items():
        print(f"{k[:60]:60s} {e['ms']:8.3f} ms  DRAM {e['traffic_bytes_per_launch'] / 1e9:7.3f} GB  algorithmic {e['algorithmic_bytes_per_launch'] / 1e9:7.3f} GB  x{e['traffic_over_algorithmic']}")


if __name__ == "__main__":
    main()
