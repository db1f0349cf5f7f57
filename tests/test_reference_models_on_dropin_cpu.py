"""The reference's OWN model files (models/resnet.py, mobilenet.py, mobile_vit.py) build and run,
unchanged, on top of this repo's drop-in `QuantizedConv2d` / `utils.quantizers`, and this repo's
workload graphs are parameter-for-parameter the same networks.  Needs the reference checkout
(/root/reference, present in the build container only) -- skipped elsewhere."""
import copy
import importlib
import os
import sys

import pytest
import torch

REF = os.environ.get("PO2_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference checkout not present")


@pytest.fixture
def ref_models_on_dropin():
    """Import the reference's `models` package with models.quantized_conv / utils.quantizers replaced
    by the drop-in shims (what INTEGRATION.md tells a user to do)."""
    import drop_in.models.quantized_conv as qc
    import drop_in.utils.quantizers as uq
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")}
    for k in saved:
        sys.modules.pop(k, None)
    sys.path.insert(0, REF)
    try:
        import models  # noqa: F401  (reference package)
        sys.modules["models.quantized_conv"] = qc
        import utils  # noqa: F401
        sys.modules["utils.quantizers"] = uq
        model_mod = importlib.import_module("models.model")
        yield model_mod
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


@pytest.mark.parametrize("name", ["resnet20", "resnet56", "mobilenet", "mobilevit"])
def test_reference_model_files_run_on_dropin_and_match_workloads(ref_models_on_dropin, name):
    import po2_quantization_b200 as P
    from workloads import mobilenet_v2_cifar, mobilevit_xs, resnet_cifar
    torch.manual_seed(8)
    ref = ref_models_on_dropin.get_model(name, 10, None, 4, (32, 32))
    qconvs = [m for m in ref.modules() if isinstance(m, P.QuantizedConv2d)]
    assert len(qconvs) == {"resnet20": 20, "resnet56": 56, "mobilenet": 50, "mobilevit": 33}[name]
    mine = {"resnet20": lambda: resnet_cifar(20), "resnet56": lambda: resnet_cifar(56),
            "mobilenet": lambda: mobilenet_v2_cifar(), "mobilevit": lambda: mobilevit_xs((32, 32), 10, (1, 1), None, 4)}[name]()
    sr, sm = ref.state_dict(), mine.state_dict()
    assert list(sr) == list(sm) and all(sr[k].shape == sm[k].shape for k in sr)
    mine.load_state_dict(sr, strict=True)
    ref.eval(); mine.eval()
    x = torch.randn(2, 3, 32, 32)
    with torch.no_grad():
        assert torch.equal(ref(x), mine(x))           # quantize_fn=None on CPU: nn.Conv2d's own path
    c = copy.deepcopy(ref)                            # test.py:120
    assert list(c.state_dict()) == list(sr)
    err, numel = ref.get_quantization_error()         # models/*.py walkers call QuantizedConv2d.get_quantization_error
    assert err == 0 and numel > 0


def test_reference_qat_model_builds_with_dropin_quantizer(ref_models_on_dropin):
    import po2_quantization_b200 as P
    m = ref_models_on_dropin.get_model("resnet20", 10, P.PowerOfTwoPlusQuantizer, 4, (32, 32))
    q = [mm for mm in m.modules() if isinstance(mm, P.QuantizedConv2d)]
    assert all(mm.quantize_fn is P.PowerOfTwoPlusQuantizer and mm.bits == 4 for mm in q)
    with pytest.raises(Exception) as ei:              # CPU tensors: loud failure, no silent fallback
        m(torch.randn(1, 3, 32, 32))
    assert "no CPU fallback" in str(ei.value)
