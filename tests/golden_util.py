"""Helpers shared by the golden-table tests (reference integration tests 1 and 2)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden():
    with open(os.path.join(HERE, "golden", "results_reference.json")) as f:
        return json.load(f)


def error_row(x, A_exact, b_exact, A, b):
    """The five printed columns of integration_test1.py:143-156 as '{:.5e}' strings."""
    dx = x[1] - x[0]
    ea = np.linalg.norm(A_exact - A, axis=0)
    eb = np.linalg.norm(b_exact - b, axis=0)
    vals = [dx, ea.max(), ea.mean(), eb.max(), eb.mean()]
    return ["{:.5e}".format(v) for v in vals]


def rows_match(got, want, last_digit_slack=1):
    """Compare '{:.5e}' strings; allow +-1 in the 6th significant digit (printing/rounding boundary)."""
    for g, w in zip(got, want):
        gm, ge = g.split("e")
        wm, we = w.split("e")
        if ge != we:
            return False
        if abs(round(float(gm) * 1e5) - round(float(wm) * 1e5)) > last_digit_slack:
            return False
    return True
