from .threshold import *  # noqa: F401,F403
