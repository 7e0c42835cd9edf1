"""GPU: the drop-in Learner as a forked process behind a real mp.Queue (boundary)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("transport", ["queue", "ring"])
def test_learner_process_end_to_end(tmp_path, transport):
    """transport="queue": the reference wire format through mp.Queue; "ring": the shared-memory
    RingQueue (same put() interface for actors, DMA straight from the shared slab)."""
    script = os.path.join(os.path.dirname(__file__), "learner_process_check.py")
    res = subprocess.run([sys.executable, script, str(tmp_path / "logs"), transport], capture_output=True,
                         text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "LEARNER_PROCESS_OK" in res.stdout
