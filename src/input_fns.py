from dalle_mtf_b200.input_fns import vae_input_fn, dalle_input_fn  # noqa: F401
