from dalle_mtf_b200.optimizers import get_optimizer, OptimizerConfig  # noqa: F401
