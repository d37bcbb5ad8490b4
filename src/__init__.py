"""Import-path compatibility with the reference tree (src/model_fns.py, src/model_fns_tf.py, src/optimizers.py,
src/input_fns.py, src/utils, src/dalle_mtf, src/vae_tf, src/data): thin re-exports of dalle_mtf_b200."""
