"""Re-export of the shared synthetic-data helpers (they live in the product package so bench.py's GPU arm
never imports oracle/)."""
from flowtron_b200.synth import *  # noqa: F401,F403
from flowtron_b200.synth import DEFAULT_MODEL_CONFIG, param_shapes, synth_params, synth_batch, beta_binomial_prior  # noqa: F401
