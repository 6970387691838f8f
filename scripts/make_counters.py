#!/usr/bin/env python3
"""profiles/r2_counters.json from ncu --set full --import-source captures: per kernel and launch, the executed
warp-instruction counts by class (source page, summed over the kernel AND the device functions it calls), DRAM bytes and
duration.  usage: python scripts/make_counters.py <n_sets> <rep> [<rep> ...] > profiles/r2_counters.json"""
import csv
import io
import json
import re
import subprocess
import sys


def source_counts(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    out, cur, hdr = {}, None, None
    launch = None
    for row in csv.reader(io.StringIO(raw)):
        if row and row[0] == "Kernel Name":
            cur = row[1]
            continue
        if row and row[0] == "Address":
            hdr = row
            continue
        if not row or hdr is None or cur is None:
            continue
        d = dict(zip(hdr, row))
        try:
            n = int(d["Instructions Executed"])
        except Exception:
            continue
        src = d["Source"].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        c = out.setdefault(cur, {"inst": 0, "imad_wide": 0, "alu": 0})
        c["inst"] += n
        if op.startswith("IMAD.WIDE"):
            c["imad_wide"] += n
        if re.match(r"(SHF|LOP3|IADD3|VIADD|SEL|PRMT|ISETP|LEA|MOV|IABS|IMNMX|VIMNMX|PLOP3|BMSK|SGXT|FLO|POPC)", op):
            c["alu"] += n
    return out


def raw_metrics(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, rows[1]))

        def val(k, to_bytes=False):
            if k not in d:
                return None
            v = float(d[k].replace(",", ""))
            if to_bytes:
                v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[k], 1)
            return v
        res.append({"name": d["Kernel Name"], "ms": val("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "s": 1e3}.get(u["gpu__time_duration.sum"].split()[0] if u.get("gpu__time_duration.sum") else "ms", 1) if val("gpu__time_duration.sum") else None,
                    "dram": (val("dram__bytes_read.sum", True) or 0) + (val("dram__bytes_write.sum", True) or 0),
                    "fmaheavy_pct": val("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                    "alu_pct": val("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active")})
    return res


def main():
    n_sets = int(sys.argv[1])
    out = {}
    for rep in sys.argv[2:]:
        sc = source_counts(rep)
        # the source page lists the kernel and every device function it calls under their own names; attribute the
        # callee counts to the (single) __global__ kernel of the report
        rm = raw_metrics(rep)
        for k in rm:
            short = re.sub(r"^void |lhb200::|bls::|mc::|<.*|\(.*", "", k["name"]).strip()
            tot = {"inst": 0, "imad_wide": 0, "alu": 0}
            for v in sc.values():
                for kk in tot:
                    tot[kk] += v[kk]
            path = "tree" if "validator" in short or "merkle" in short or "hash" == short[:4] and "g2" not in short else "bls"
            out[short] = {"path": path, "n_sets": n_sets if path == "bls" else None, "ms_under_ncu": k["ms"],
                          "warp_inst": tot["inst"], "imad_wide_warp_inst": tot["imad_wide"], "alu_warp_inst": tot["alu"],
                          "dram_bytes": k["dram"], "pipe_fmaheavy_pct": k["fmaheavy_pct"], "pipe_alu_pct": k["alu_pct"],
                          "report": rep.split("/")[-1]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
