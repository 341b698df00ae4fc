"""Host logic of the drop-in (argument parsing, state machine, iterator/set() bookkeeping,
value stores, flatten + filter construction) replayed over the committed golden vectors.

No GPU here: Automaton._scan_flat is routed through tests/emul.py, a pure-Python restatement
of the two device kernels operating on the tables the host core produced.  The same scenarios
run against the real kernels in tests/test_gpu_parity.py (-m gpu).
"""
import pytest

import emul
import pyahocorasick_b200 as ac
from golden_driver import all_scenarios, run_ops

SC = all_scenarios()


@pytest.mark.parametrize("algo", ["filter", "dfa"])
@pytest.mark.parametrize("sc", SC, ids=[s["name"] for s in SC])
def test_golden_via_emulated_device(sc, algo, monkeypatch):
    emul.install(monkeypatch, algo)
    bad = run_ops(ac.flavour(sc["flavour"]), sc, record=False)
    assert not bad, bad[:3]
