from synthetic_env import SyntheticOfflineEnv as Env  # noqa: F401
from synthetic_env import make  # noqa: F401
