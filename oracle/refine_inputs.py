"""Seeded refiner inputs + the REGRESSION sections of the shipped refiner configs
(refining/tools/cfgs/ref_model_cfgs/vehicle_{prm,grm,crm}_model.yaml) -- TEST INFRASTRUCTURE."""
from detzero_b200.synthetic import (prm_cfg, grm_cfg, crm_cfg, prm_inputs, grm_inputs, crm_inputs)    # noqa: F401  (one definition)
