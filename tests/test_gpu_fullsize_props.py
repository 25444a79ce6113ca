"""-m gpu: the size-independent property checks of tools/fullsize_check.py at a moderate scale
(the tool itself is run at BASELINE's full sizes on the GPU box; results under profiles/)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("workload,scale", [("config5", 0.01), ("config3", 0.02), ("config4", 0.002)])
def test_properties_and_oracle_spot_check(workload, scale):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fullsize_check.py"), "--workload", workload, "--scale", str(scale),
                        "--spot", "60"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["signatures"] > 100 and out["clusters"] > 10 and out["oracle_spot_checked_partitions"] > 10
    if workload == "config5":
        assert out["partitions_over_100"] >= 1          # the host RNG sampling path is exercised
