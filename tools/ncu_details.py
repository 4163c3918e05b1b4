"""ncu report -> the plain-text details table kept under profiles/ (one line per metric).  usage: ncu_details.py rep out.txt"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "details", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
with open(sys.argv[2], "w") as f:
    for r in rows[1:]:
        r = r + [""] * (len(hdr) - len(r))
        if not r[ix["Metric Name"]]:
            continue
        f.write(" | ".join([r[ix["ID"]], r[ix["Kernel Name"]].replace("vp::", ""), r[ix["Section Name"]], r[ix["Metric Name"]], r[ix["Metric Unit"]], r[ix["Metric Value"]]]) + "\n")
