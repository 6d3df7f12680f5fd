class EnvSpec:
    pass
