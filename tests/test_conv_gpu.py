"""GPU checks of the grouped tcgen05 conv / linear kernels against exact-fp32 PyTorch (fprop, dgrad, wgrad+SGD)."""
import pytest
import torch

from murmura_b200.ops import selfcheck as sc

pytestmark = pytest.mark.gpu

SMALL = [c for c in sc.CASES if not c[0].startswith(("rn.conv1", "leaf.fc1"))]


@pytest.fixture(scope="module")
def harness():
    return sc.Harness(torch.device("cuda", 0))


@pytest.mark.parametrize("mode", ["F", "D", "W"])
@pytest.mark.parametrize("case", SMALL, ids=[c[0] for c in SMALL])
def test_conv_gemm_matches_fp32_reference(harness, case, mode):
    r = sc.check_case(harness, case, mode)
    assert r.get("skipped") or r["ok"], r


@pytest.mark.parametrize("mode", ["F", "D", "W"])
@pytest.mark.parametrize("G,splitk,perm", [(3, 1, True), (1, 3, False), (2, 4, True)])
def test_grouped_and_split_k(harness, mode, G, splitk, perm):
    for name in ("tiny3x3", "rn.layer3", "har.fc1"):
        case = next(c for c in sc.CASES if c[0] == name)
        r = sc.check_case(harness, case, mode, G=G, splitk=splitk, perm_groups=perm)
        assert r["ok"], r


@pytest.mark.parametrize("mode", ["F", "W"])
def test_first_layer_of_resnet(harness, mode):
    case = next(c for c in sc.CASES if c[0] == "rn.conv1")
    r = sc.check_case(harness, case, mode)
    assert r["ok"], r
