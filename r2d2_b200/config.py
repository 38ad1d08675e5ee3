"""Module-level constants with the reference's names and defaults (config.py:1-37 upstream).

They are read both as def-time default arguments and at run time (training_steps,
learning_starts, log_interval, learning_steps, block_length, forward_steps), exactly like
the reference, so `import config; config.x = ...` before constructing the workers behaves
the same way.

Upstream the only way to change a value is to edit config.py before starting; the equivalent here, without editing the
package, is the environment variable R2D2_CONFIG_OVERRIDES holding a JSON object of {name: value} that is applied when
this module is first imported (i.e. before any worker captures a def-time default), e.g.
    R2D2_CONFIG_OVERRIDES='{"training_steps": 1000, "num_actors": 4}' python train.py
Unknown names are an error (a typo must not silently train with the defaults).
"""
game_name = 'MsPacman'
obs_shape = (1, 84, 84)

lr = 1e-4
eps = 1e-3
grad_norm = 40
batch_size = 64
learning_starts = 50000
save_interval = 500
target_net_update_interval = 2000
gamma = 0.997
prio_exponent = 0.9
importance_sampling_exponent = 0.6

training_steps = 100000
buffer_capacity = 2000000
max_episode_steps = 27000
actor_update_interval = 400
block_length = 400  # cut one episode to numbers of blocks to improve the buffer space utilization

num_actors = 8
base_eps = 0.4
alpha = 7
log_interval = 10

# sequence setting
burn_in_steps = 40
learning_steps = 40
forward_steps = 5
seq_len = burn_in_steps + learning_steps + forward_steps

# network setting
hidden_dim = 512

render = False
save_plot = True
test_epsilon = 0.001


def _apply_overrides():
    import json
    import os
    raw = os.environ.get("R2D2_CONFIG_OVERRIDES")
    if not raw:
        return
    g = globals()
    for name, value in json.loads(raw).items():
        if name not in g or name.startswith("_"):
            raise KeyError(f"R2D2_CONFIG_OVERRIDES: unknown config name {name!r}")
        g[name] = tuple(value) if isinstance(g[name], tuple) else value
    g["seq_len"] = g["burn_in_steps"] + g["learning_steps"] + g["forward_steps"]


_apply_overrides()
