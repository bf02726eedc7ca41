from .matcher import *   # noqa: F401,F403  (reference: `from .matcher import *`, __all__ = ['Matcher'])
