"""Oracle vs the reference's own code on FRESH random inputs (tests/golden/live_check.py), in a subprocess so that
the shim and the reference's top-level packages never enter this test session.  Needs /root/reference: runs in the
build container, skipped on the GPU box (the committed fixtures are what travels)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('RON_REFERENCE', '/root/reference')


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'nets')), reason='the reference tree is not on this machine')
@pytest.mark.parametrize('seed', [20261, 20262])
def test_oracle_equals_reference_on_fresh_inputs(seed):
    r = subprocess.run([sys.executable, os.path.join(HERE, 'golden', 'live_check.py'), str(seed), '3'],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'live check ok' in r.stdout
