"""Host-side mirror of src/optimizers.py: hyper-parameter defaults, learning-rate schedule, optimiser wiring.

The schedule is a handful of scalar flops per step and is evaluated on the host from the integer step counter (no
device sync); clip-by-global-norm and Adam themselves are the K9/K10 kernels (csrc/optim.cu).
"""
import math


class OptimizerConfig:
    """Defaults exactly as src/optimizers.py:24-28, 84-87."""

    def __init__(self, params):
        self.lr = params["lr"]
        self.end_step = params.get("lr_decay_end") or params["train_steps"]   # optimizers.py:24
        self.lr_decay = params.get("lr_decay") or "cosine"                    # optimizers.py:25
        ws = params.get("warmup_steps")
        self.warmup_steps = 3000 if ws is None else ws                        # optimizers.py:26
        gc = params.get("gradient_clipping")
        self.gradient_clipping = 1.0 if gc is None else gc                    # optimizers.py:27
        self.optimizer = (params.get("optimizer") or "adam").lower()          # optimizers.py:28
        self.weight_decay = params.get("weight_decay") or 0.0                 # optimizers.py:84
        self.beta_1 = params.get("beta_1") or 0.9
        self.beta_2 = params.get("beta_2") or 0.999
        self.epsilon = params.get("epsilon") or 1e-6
        if self.optimizer != "adam":
            # the Adafactor branch (optimizers.py:90-97) is selected by no reference config
            raise ValueError(f"{self.optimizer} not recognized (only 'adam' is implemented on B200)")

    def learning_rate(self, step):
        """Value used for the update taken at global step `step` (the step counter before the increment,
        src/model_fns.py:201).  cosine: tf.train.cosine_decay(alpha=0.1); linear: polynomial_decay to 0.1*lr;
        then linear warm-up (optimizers.py:46-76)."""
        s = min(step, self.end_step)
        if self.lr_decay == "linear":
            lr = (self.lr - 0.1 * self.lr) * (1.0 - s / self.end_step) + 0.1 * self.lr
        elif self.lr_decay == "cosine":
            lr = self.lr * (0.9 * 0.5 * (1.0 + math.cos(math.pi * s / self.end_step)) + 0.1)
        else:
            lr = self.lr
        if self.warmup_steps > 0 and step < self.warmup_steps:
            lr = lr * (step / self.warmup_steps)
        return lr


def get_optimizer(engine, params):
    """Mirror of get_optimizer(mesh, loss, params, variable_dtype) (src/optimizers.py:19): returns
    (learning_rate_fn, update_op) where update_op(step) applies clip + Adam to `engine` on the device."""
    cfg = OptimizerConfig(params)

    def update_op(step):
        lr = cfg.learning_rate(step)
        engine.optimizer_step(lr, beta1=cfg.beta_1, beta2=cfg.beta_2, eps=cfg.epsilon,
                              weight_decay=cfg.weight_decay, clip=cfg.gradient_clipping)
        return lr

    return cfg.learning_rate, update_op
