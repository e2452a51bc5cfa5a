"""`-d` colour classes (wgdi/dotplot.py:31-56): the oracle and the product's data function against
what the unmodified reference's `pair_positon` returned (tests/make_golden.py colors -> e2e_<case>/expected.colors.npz).
The per-query sort is numpy's unstable argsort (hazard H2), so fixtures were written, and are replayed, in a process with
numpy's SIMD sorts disabled; the child checks the argsort fingerprint first."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import PORTABLE_ENV, ROOT, stage_e2e

_CHILD = r"""
import sys, hashlib, json, numpy as np
sys.path.insert(0, %r)
fp = hashlib.md5(np.argsort((np.arange(1000) * 7919 %% 5).astype(np.int64), kind="quicksort").tobytes()).hexdigest()
if fp != json.load(open("meta.json"))["argsort_fingerprint"]:
    print("SKIP fingerprint", fp); sys.exit(0)
from oracle import icl_oracle
from wgdi_b200 import base
from wgdi_b200.dotplot import dotplot
gold = np.load("expected.colors.npz")
icl = dict(base.load_conf("icl.conf", "collinearity"))
sets = sorted({k.rsplit("_", 1)[0] for k in gold.files})
assert len(sets) == 12
drawn = 0
for key in sets:
    reverse, mult, rn = key.split("_")
    exp_row, exp_color = gold[key + "_row"], gold[key + "_color"]
    for ingest in ("native", "pandas"):
        for blast in ("blast.m8", "blast.m8.gz"):
            opts = [(k, icl[k]) for k in ("gff1", "gff2", "lens1", "lens2")] + [
                ("blast", blast), ("blast_reverse", reverse), ("multiple", mult), ("repeat_number", rn), ("score", "100"),
                ("evalue", "1e-5"), ("position", "order"), ("ingest", ingest), ("markersize", "0.5"), ("savefig", "x.png")]
            got = dotplot(opts).colored_hits()
            assert (got["row"] == exp_row).all() and (got["color"] == exp_color).all(), (key, ingest, blast)
            assert (got["color_name"] == np.array(["", "red", "blue", "gray"], dtype=object)[exp_color]).all()
    # the oracle on the filtered table (the pandas reader keeps the reference's row labels)
    gff1, gff2 = base.newgff(icl["gff1"]), base.newgff(icl["gff2"])
    fr = base.newblast("blast.m8", 100, 1e-5, gff1, gff2, reverse)
    oc = icl_oracle.color_hits(fr["g1"].to_numpy(), fr[11].to_numpy(), int(mult), int(rn))
    assert (fr.index.to_numpy()[oc > 0] == exp_row).all() and (oc[oc > 0] == exp_color).all(), key
    drawn += len(exp_row)
print("OK", drawn)
"""


@pytest.mark.parametrize("case", ["tiny", "cross"])
def test_colour_classes_match_the_reference(case):
    with tempfile.TemporaryDirectory() as d:
        stage_e2e(case, d)
        env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=PORTABLE_ENV)
        r = subprocess.run([sys.executable, "-c", _CHILD % ROOT], cwd=d, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        if r.stdout.startswith("SKIP"):
            pytest.skip("numpy scalar-sort fingerprint differs on this machine: " + r.stdout)
        assert r.stdout.startswith("OK") and int(r.stdout.split()[1]) > 1000, r.stdout


def test_color_hits_equals_oracle_on_ties():
    """Tie-heavy random tables in THIS process (whatever numpy's argsort flavour is here): vectorised rank classes ==
    the oracle's per-query loop, including multiple = 0, blue wider than repeat_number and float scores."""
    from oracle import icl_oracle
    from wgdi_b200.ingest import color_hits
    rng = np.random.default_rng(23)
    for trial in range(150):
        n, nq = int(rng.integers(1, 500)), int(rng.integers(1, 30))
        q = rng.integers(0, nq, n)
        bits = rng.integers(100, 104 + trial % 7 * 20, n).astype(np.int64 if trial % 3 else np.float64)
        mult, rn = int(rng.integers(0, 5)), int(rng.integers(1, 30))
        assert (color_hits(q, bits, mult, rn) == icl_oracle.color_hits(q, bits, mult, rn)).all(), trial
