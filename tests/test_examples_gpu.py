"""The "one node serves B drones per step" example (examples/serve_drones.py): signed state
messages in, mission goals, ONE batched solve, signed trajectory messages out."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_serve_drones_example(oracle_mod):
    spec = importlib.util.spec_from_file_location("serve_drones", os.path.join(ROOT, "examples", "serve_drones.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sol, outbox, stats = mod.serve(B=192, steps=4, verbose=False)
    assert len(outbox) == 192 and len(stats) == 4
    assert sol.x.shape[0] == 192 and (sol.status >= 0).all()
    from dart_planner_b200.wire import SignedEnvelope, trajectory_from_payload
    tr = trajectory_from_payload(SignedEnvelope(secret_key="demo-secret").deserialize(outbox[7]))
    np.testing.assert_array_equal(tr.positions, sol.positions[7])
