"""Pack the instrument transmission tables into one binary resource.

The 18 response tables and the wheel specification that ship with
aconley/mbb_emcee (mbb_emcee/resources/*.txt) are
*measurements*, not code.  The product needs them as data (the default
``response_set()`` of the reference loads them, response.py:680-686), but the
reference tree does not exist on the GPU box, so they are re-shipped here in a
single compressed ``.npz``:

    wheel_names   [18]   response names          (mbb_filterwheel.txt col 0)
    wheel_files   [18]   original table name     (col 1; kept for provenance)
    wheel_xtype / wheel_xunits / wheel_senstype / wheel_normtype  [18] (cols 2-5)
    wheel_xnorm / wheel_normparam                [18] float64 (cols 6-7)
    x_<i>, r_<i>                                 raw float64 columns of table i

Values are the correctly-rounded float64 of each text token, i.e. bit-for-bit
what the reference reads.  Run:  python tools/pack_filters.py [refdir]
"""
import os
import sys

import numpy as np


def read_table(path):
    rows = []
    with open(path) as fh:
        for line in fh:
            s = line.strip()
            if not s or s.startswith("#"):
                continue
            rows.append(s.split())
    return rows


def main(refdir="/root/reference"):
    res = os.path.join(refdir, "mbb_emcee", "resources")
    spec = read_table(os.path.join(res, "mbb_filterwheel.txt"))
    out = {
        "wheel_names": np.array([r[0] for r in spec]),
        "wheel_files": np.array([r[1] for r in spec]),
        "wheel_xtype": np.array([r[2].lower() for r in spec]),
        "wheel_xunits": np.array([r[3].lower() for r in spec]),
        "wheel_senstype": np.array([r[4].lower() for r in spec]),
        "wheel_normtype": np.array([r[5].lower() for r in spec]),
        "wheel_xnorm": np.array([float(r[6]) for r in spec]),
        "wheel_normparam": np.array([float(r[7]) for r in spec]),
    }
    nn = 0
    for i, r in enumerate(spec):
        tab = read_table(os.path.join(res, r[1]))
        out["x_%d" % i] = np.array([float(t[0]) for t in tab], dtype=np.float64)
        out["r_%d" % i] = np.array([float(t[1]) for t in tab], dtype=np.float64)
        nn += len(tab)
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "mbb_emcee_b200", "resources", "filterwheel.npz")
    np.savez_compressed(dst, **out)
    print("packed %d filters, %d nodes -> %s (%d bytes)" %
          (len(spec), nn, os.path.normpath(dst), os.path.getsize(dst)))


if __name__ == "__main__":
    main(*sys.argv[1:])
