"""Reference-script-style analysis scenarios (tests/scenarios.py): two-stage linear buckling, pre-stressed natural
frequencies, buckling under a given stress, consistent / lumped mass frequencies of the line elements, material
coordinates -- physics scalars from (a) the numpy oracle [CPU] and (b) the CUDA path [gpu] against the values obtained
with the compiled reference (tests/golden/scenario_scalars.json) at north_star's 1e-8, plus the analytic checks the
reference's own tests make (tests/test_beamlr_natural_freq_cantilever.py:106-111, tests/test_truss_natural_freq.py:98-104)."""
import json
import os

import pytest

from oracle import driver
from tests import scenarios, util

GOLD = json.load(open(os.path.join(util.GOLDEN_DIR, "scenario_scalars.json")))
RTOL = 1e-8


def _compare(name, got):
    assert set(got) == set(GOLD[name])
    for k, want in GOLD[name].items():
        assert abs(got[k] - want) <= RTOL * abs(want), (name, k, got[k], want)


def test_reference_scalars_satisfy_the_references_analytic_checks():
    assert abs(GOLD["beamlr_cantilever_freq"]["omega1_over_euler"] - 1.) <= 0.015
    assert abs(GOLD["truss_freq"]["omega1_over_exact"] - 1.) <= 0.01
    assert GOLD["beamc_prestress_freq"]["omega1_prestress"] < GOLD["beamc_prestress_freq"]["omega1"]     # compression softens
    assert GOLD["quad4r_prestress_freq"]["omega1_prestress"] > GOLD["quad4r_prestress_freq"]["omega1"]   # tension stiffens


@pytest.mark.parametrize("name", scenarios.NAMES)
def test_oracle_reproduces_reference_scenarios(name):
    _compare(name, scenarios.SCENARIOS[name](scenarios.evaluate_with(driver.run)))


@pytest.mark.gpu
@pytest.mark.parametrize("name", scenarios.NAMES)
def test_cuda_reproduces_reference_scenarios(name):
    _compare(name, scenarios.SCENARIOS[name](scenarios.evaluate_cuda))
