"""Small host-side helpers shared by the API layer."""
from collections.abc import Iterable

import numpy

__all__ = ["isiterable"]


def isiterable(obj):
    """True when ``obj`` can be looped over (0-d arrays cannot).

    Mirrors reference mbb_emcee/utility.py:7-21; the scalar/array distinction
    decides which evaluation path ``modified_blackbody.__call__`` takes
    (reference modified_blackbody.py:549-554), so it must agree exactly.
    """
    if isinstance(obj, numpy.ndarray):
        return obj.ndim != 0
    if isinstance(obj, Iterable):
        return True
    try:
        iter(obj)
    except TypeError:
        return False
    return True


def read_text_table(path):
    """Rows of whitespace-separated tokens; blank and '#' lines skipped.

    Numeric tokens become float, anything else stays a string.  This is the
    only behaviour the reference needs from ``astropy.io.ascii.read``
    (response.py:198-203, 699-710; likelihood.py:254-262); astropy is not
    required.
    """
    rows = []
    with open(path, "r") as handle:
        for raw in handle:
            line = raw.strip()
            if not line or line[0] == "#":
                continue
            row = []
            for tok in line.split():
                try:
                    row.append(float(tok))
                except ValueError:
                    row.append(tok)
            rows.append(tuple(row))
    return rows
