"""Auto-skipping live cross-check: where a real pymunk < 6 and a checkout of the reference
are available, the committed golden fixtures -- generated over the restated Chipmunk in oracle/shims -- are
re-generated with the real library and must agree.  Skipped in the build image and on the GPU box (no pymunk)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def test_golden_fixtures_agree_with_live_pymunk():
    import live_check
    pm, why = live_check.real_pymunk()
    if pm is None:
        pytest.skip("no real pymunk < 6 here (%s): Chipmunk layer stays 'parity unpinned'" % why)
    if live_check.reference_path() is None:
        pytest.skip("reference checkout not found (set SHIPSIM_REF_PATH)")
    n, steps, worst = live_check.run(verbose=False)
    assert n > 30 and steps > 2000


def test_live_check_refuses_the_shim():
    """The hook must never mistake oracle/shims/pymunk for the real library."""
    import live_check
    shim = os.path.join(ROOT, "oracle", "shims")
    sys.path.insert(0, shim)
    try:
        pm, why = live_check.real_pymunk()
    finally:
        if shim in sys.path:
            sys.path.remove(shim)
        sys.modules.pop("pymunk", None)
    assert pm is None or "minimunk" not in str(getattr(pm, "version", ""))
