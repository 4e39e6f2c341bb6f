"""python-decouple stand-in (natrix/core/common/constants.py:3): environment, then the default."""
import os


def config(name, default=None, cast=None):
    value = os.environ.get(name, default)
    return cast(value) if cast is not None and value is not None else value
