#!/usr/bin/env python
"""Convert the reference's Desmond .dms example systems into compact .npz fixtures (positions + AGBNP parameters only).

Run in the build container (where /root/reference exists); the .npz files are committed under tests/golden/systems/
because /root/reference does not exist on the GPU box.  Loader rule (the app-layer rule of
OpenMM's DesmondDMSFile.createSystem(implicitSolvent='AGBNP') is not part of the reference repo):
  SELECT particle.id, anum, x, y, z, charge, radius, igamma, ialpha, salpha FROM particle JOIN agbnp2 USING(id)
  x,y,z,radius: Angstrom -> nm (x0.1);  gamma = igamma*4.184/0.01 kJ/mol/nm^2;
  alpha = (ialpha+salpha)*4.184e-3... see below;  ishydrogen = (anum == 1)
Stored arrays are float64 in file units converted to OpenMM units (nm, kJ/mol, e).
"""
import os
import sqlite3
import sys
import numpy as np

REF_EXAMPLES = "/root/reference/example"
SYSTEMS = ["trpcage_agbnp1", "1li2_agbnp1", "rnaseh_agbnp1", "1dwc_agbnp1", "2clr_agbnp1", "hivrt_agbnp1"]

ANG2NM = 0.1
KCAL2KJ = 4.184


def load_dms(path):
    con = sqlite3.connect("file:%s?mode=ro" % path, uri=True)
    rows = con.execute(
        "SELECT particle.id, anum, x, y, z, charge, radius, igamma, ialpha, salpha "
        "FROM particle JOIN agbnp2 USING(id) ORDER BY particle.id").fetchall()
    con.close()
    a = np.array(rows, dtype=np.float64)
    anum = a[:, 1].astype(np.int32)
    pos = a[:, 2:5] * ANG2NM
    charge = a[:, 5].copy()
    radius = a[:, 6] * ANG2NM
    gamma = a[:, 7] * KCAL2KJ / (ANG2NM * ANG2NM)
    # alpha is in kcal/mol A^3 in the file -> kJ/mol nm^3
    alpha = (a[:, 8] + a[:, 9]) * KCAL2KJ * ANG2NM ** 3
    ish = (anum == 1).astype(np.int32)
    return dict(pos=pos, radius=radius, gamma=gamma, alpha=alpha, charge=charge, ishydrogen=ish)


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "systems")
    os.makedirs(out_dir, exist_ok=True)
    for name in SYSTEMS:
        path = os.path.join(REF_EXAMPLES, name + ".dms")
        if not os.path.exists(path):
            print("skip (missing):", path)
            continue
        s = load_dms(path)
        short = name.replace("_agbnp1", "")
        np.savez_compressed(os.path.join(out_dir, short + ".npz"), **s)
        print(short, "N =", len(s["radius"]), "heavy =", int((s["ishydrogen"] == 0).sum()),
              "radii(A) =", sorted(set(np.round(s["radius"] * 10, 4))))


if __name__ == "__main__":
    main()
