"""Drop-in mounting: ``breaching_b200.install.install()`` rebinds ``breaching.attacks.prepare_attack`` of the (unmodified)
reference package, so reference entry points keep calling ``breaching.attacks.prepare_attack(...)`` unchanged.
Needs the reference tree (build container only); skipped elsewhere."""
import pytest
import torch

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="reference tree not present")


def test_install_rebinds_prepare_attack_and_delegates_other_attack_types():
    ref = refshim.import_reference()
    import breaching_b200
    from breaching_b200 import install as inst
    from breaching_b200 import synthetic
    from breaching_b200.engine import EngineError

    original = ref.attacks.prepare_attack
    try:
        returned = inst.install()
        assert returned is original
        assert ref.attacks.prepare_attack.__module__.startswith("breaching_b200")
        model = synthetic.build_model("convnet-tiny", 10)
        loss = torch.nn.CrossEntropyLoss()
        setup = dict(device=torch.device("cpu"), dtype=torch.float)
        # optimisation attacks go to the B200 engine: on a CPU "device" it refuses loudly (no fallback) ...
        with pytest.raises(EngineError):
            ref.attacks.prepare_attack(model, loss, breaching_b200.get_attack_config("invertinggradients"), setup)
        # ... while the attack types outside the accelerated path are delegated to the reference's own classes
        cfg = refshim.load_reference_attack_cfg("analytic")
        attacker = ref.attacks.prepare_attack(model, loss, cfg, setup)
        assert type(attacker).__module__.startswith("breaching.attacks")
    finally:
        inst.uninstall()
    assert ref.attacks.prepare_attack is original


def test_reference_yaml_config_objects_are_accepted_by_the_engine_config_flattening():
    """cfg objects composed from the reference's own YAML (attribute + item access) flatten to the same C struct as ours."""
    import ctypes

    import breaching_b200
    from breaching_b200.engine import make_cfg

    for name in ["invertinggradients", "modern", "seethroughgradients", "clsattack", "legacy"]:
        a = make_cfg(refshim.load_reference_attack_cfg(name))
        b = make_cfg(breaching_b200.get_attack_config(name))
        assert bytes(ctypes.string_at(ctypes.addressof(a), ctypes.sizeof(a))) == bytes(ctypes.string_at(ctypes.addressof(b), ctypes.sizeof(b))), name
