"""profiles/ncu_counts.json from `ncu --page raw --csv` exports of one k_map and one k_fir launch of the bench batch:
python tools/ncu_counts.py raw_k_map.csv raw_k_fir.csv "<source note>"   (bench.py uses the counts only while csrc/ is unchanged)"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402


def one(path, pattern):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    for vals in rows[2:]:
        if pattern in vals[hdr.index("Kernel Name")]:
            g = lambda k: float(vals[hdr.index(k)].replace(",", ""))
            unit = lambda k: rows[1][hdr.index(k)]
            def bytes_of(k):
                u = unit(k)
                return g(k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            return {"inst_executed": g("smsp__inst_executed.sum"), "dram_bytes": bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum"),
                    "time_us_under_ncu": g("gpu__time_duration.sum") * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(unit("gpu__time_duration.sum"), 1)}
    raise SystemExit("no %s launch in %s" % (pattern, path))


out = {"csrc_sha256": kernel_source_hash(), "source": sys.argv[3] if len(sys.argv) > 3 else "", "k_map": one(sys.argv[1], "k_map"),
       "k_fir": one(sys.argv[2], "k_fir")}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_counts.json"), "w"), indent=1)
print(out)
