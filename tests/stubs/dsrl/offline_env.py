def wrap_env(env, reward_scale=None, **k):
    return env


def OfflineEnvWrapper(env):
    return env
