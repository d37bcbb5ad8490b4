from dalle_mtf_b200.model_fns import dalle_model_fn, load_vae_model, mode_to_str, StepSpec, TRAIN, EVAL, PREDICT  # noqa: F401
