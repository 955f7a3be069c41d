"""Input systems for tests and the benchmark: committed fixtures derived from the reference's example .dms files
(tools/dms_to_npz.py), a direct .dms reader, and the deterministic HIV-RT stand-in."""
import os
import sqlite3

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE_DIR = os.path.join(_ROOT, "tests", "golden", "systems")

ANG2NM = 0.1
KCAL2KJ = 4.184


def load_dms(path):
    """Desmond .dms (SQLite) -> AGBNP parameters in OpenMM units (nm, kJ/mol, e).  Loader rule: SURVEY.md 8d."""
    con = sqlite3.connect("file:%s?mode=ro" % path, uri=True)
    rows = con.execute(
        "SELECT particle.id, anum, x, y, z, charge, radius, igamma, ialpha, salpha "
        "FROM particle JOIN agbnp2 USING(id) ORDER BY particle.id").fetchall()
    con.close()
    a = np.array(rows, dtype=np.float64)
    return dict(pos=a[:, 2:5] * ANG2NM, charge=a[:, 5].copy(), radius=a[:, 6] * ANG2NM,
                gamma=a[:, 7] * KCAL2KJ / (ANG2NM * ANG2NM), alpha=(a[:, 8] + a[:, 9]) * KCAL2KJ * ANG2NM ** 3,
                ishydrogen=(a[:, 1].astype(np.int32) == 1).astype(np.int32))


def load(name):
    """Load a committed fixture (trpcage, 1li2, rnaseh, 1dwc, 2clr) or the HIV-RT input."""
    if name in ("hivrt", "hivrt_standin"):
        return hivrt()
    s = np.load(os.path.join(FIXTURE_DIR, name + ".npz"))
    return {k: s[k] for k in s.files}


def float_rounded(pos):
    """Positions rounded to float once; the same values go to the oracle and to the GPU."""
    return np.asarray(pos, dtype=np.float32).astype(np.float64)


def hivrt():
    """example/hivrt_agbnp1.dms if the driver supplied it, else the deterministic stand-in "2clr x 3" (N = 17 949):
    three copies of 2clr translated by (0,0,0), (+5.5,0,0), (0,0,+6.0) nm."""
    real = os.path.join(_ROOT, "example", "hivrt_agbnp1.dms")
    if os.path.exists(real):
        s = load_dms(real)
        s["name"] = "hivrt_agbnp1.dms"
        return s
    b = load("2clr")
    shifts = np.array([[0.0, 0.0, 0.0], [5.5, 0.0, 0.0], [0.0, 0.0, 6.0]])
    base = float_rounded(b["pos"])
    out = {k: np.concatenate([b[k]] * 3) for k in ("radius", "gamma", "alpha", "charge", "ishydrogen")}
    out["pos"] = np.concatenate([base + sh for sh in shifts])
    out["name"] = "hivrt-standin-2clr-x3"
    return out


def jitter(pos, seed, amplitude=0.001):
    """Seeded uniform +-amplitude nm displacement of every coordinate (forces a genuine tree rebuild, as MD would)."""
    rng = np.random.default_rng(seed)
    return pos + rng.uniform(-amplitude, amplitude, size=pos.shape)


def make_force(s, version=1, method=0, cutoff=1.0):
    from .AGBNPplugin import AGBNPForce
    f = AGBNPForce()
    f.setVersion(version)
    f.setNonbondedMethod(method)
    f.setCutoffDistance(cutoff)
    for i in range(len(s["radius"])):
        f.addParticle(s["radius"][i], s["gamma"][i], s["alpha"][i], s["charge"][i], bool(s["ishydrogen"][i]))
    return f
