from dalle_mtf_b200.model_fns import vae_model_fn, vae_temperature, TRAIN, EVAL, PREDICT  # noqa: F401
