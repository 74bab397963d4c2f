"""Drop-in for the reference's ``model.main_model.model`` module: set ``model.architecture:
"playablevideogeneration_b200.model.main_model.model"`` in the YAML and train.py:38-39 / play.py:45-46 pick it up."""
from playablevideogeneration_b200.caddy import Model  # noqa: F401


def model(config):
    return Model(config, reduced=False)
