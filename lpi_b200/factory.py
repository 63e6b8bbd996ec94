"""utils/factory.py:3-7 of the reference."""
from .sprompt import SPrompts


def get_model(model_name, args):
    options = {"sprompts": SPrompts}
    return options[model_name.lower()](args)
