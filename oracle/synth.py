"""TEST INFRASTRUCTURE re-export: the synthetic weight / batch generators live in the product package
(`pose_adv_aug_b200/synth.py`, plain host code used by bench.py and smoke()); the golden generators, the
oracle tests and the GPU parity tests keep importing them from here."""
from pose_adv_aug_b200.synth import _rng, schema_of, make_state_dict, make_images, make_heatmaps, make_tensor, make_photo   # noqa: F401
