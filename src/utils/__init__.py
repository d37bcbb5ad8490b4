from dalle_mtf_b200.utils import (fetch_model_params, yes_or_no, remove_gs_or_filepath, maybe_remove_gs_or_filepath,  # noqa: F401
                                  setup_logging, print_n_params, parse_mesh, local_path, latest_checkpoint,
                                  load_global_step_from_checkpoint_dir, save_checkpoint, load_checkpoint)
from dalle_mtf_b200.model_fns import mode_to_str  # noqa: F401
