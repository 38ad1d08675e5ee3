"""Drop-in alias: with this directory on PYTHONPATH the reference's train.py (`import model` /
`from model import ...`) binds the B200 implementation -- this name IS r2d2_b200.model, so run-time edits
such as `config.training_steps = N` reach the workers exactly as they do upstream."""
import sys

import r2d2_b200.model as _impl

sys.modules[__name__] = _impl
