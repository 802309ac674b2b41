"""-m gpu: the CUDA path through the C-ABI against the rows the REAL reference produced."""
import pytest

import golden_util
from oracle import compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_util.case_names())
def test_cuda_matches_reference_rows(engine, name):
    case = golden_util.load_case(name)
    engine.set_params(case["params"])
    engine.set_contigs(case["lens"])
    res = engine.cluster(case["sigs"], case["reads"])
    d = compare.diff_rows(case["rows"], golden_util.to_rows(case, res))
    assert not d, "\n".join(d)
