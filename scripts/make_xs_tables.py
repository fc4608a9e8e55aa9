"""Pack the reference's cross-section tables (monte_cpp/xcom2.csv, Ca.csv) into
monte_b200/data/xs_tables.npz so tests and bench.py have them on boxes without /root/reference.

Parsing follows readcsv (CBCT_real2.cpp:633-668): 200 rows "coh,compton,photo,total", row r is
index r+1 = keV; UTF-8 BOM and CRLF are stripped; index 0 duplicates index 1.  Values are stored
as float64 exactly as strtod() reads the tokens.  The BOM work-around csvarray[0][1]=1.372
(CBCT_real2.cpp:663) is NOT baked in; loaders apply it on request.
    python scripts/make_xs_tables.py
"""
import os
import sys

import numpy as np

REF = "/root/reference/monte_cpp"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "monte_b200", "data", "xs_tables.npz")


def parse(path):
    t = np.zeros((4, 201), np.float64)
    with open(path, "rb") as f:
        raw = f.read()
    if raw.startswith(b"\xef\xbb\xbf"):
        raw = raw[3:]
    rows = [ln for ln in raw.decode("ascii").replace("\r", "").split("\n") if ln.strip()]
    assert len(rows) == 200, len(rows)
    for r, ln in enumerate(rows):
        t[:, r + 1] = [float(x) for x in ln.split(",")]
    t[:, 0] = t[:, 1]
    return t


if __name__ == "__main__":
    h2o, ca = parse(os.path.join(REF, "xcom2.csv")), parse(os.path.join(REF, "Ca.csv"))
    # known answers, SURVEY.md 8a-A1
    assert abs(h2o[3, 140] - 0.1538092) < 1e-12 and abs(h2o[3, 60] - 0.20585) < 1e-5 and abs(ca[3, 140] - 0.17730) < 1e-12
    np.savez_compressed(OUT, h2o=h2o, ca=ca, columns=np.array(["coh", "compton", "photo", "total"]),
                        density_h2o=1.0, density_ca=1.55)
    print(OUT, os.path.getsize(OUT))
