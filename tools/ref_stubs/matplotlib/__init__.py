"""Import stub used only by tools/gen_golden*.py to import the reference in this container."""
__version__ = '0.0-stub'
from . import path, colors  # noqa
