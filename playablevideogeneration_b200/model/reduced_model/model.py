"""Drop-in for the reference's ``model.reduced_model.model`` (Breakout, configs/02_breakout.yaml:25)."""
from playablevideogeneration_b200.caddy import Model as _Model


class Model(_Model):
    def __init__(self, config):
        super().__init__(config, reduced=True)


def model(config):
    return Model(config)
