class VectorEnv:
    pass


class AutoresetMode:
    NEXT_STEP = "next_step"
    SAME_STEP = "same_step"
    DISABLED = "disabled"
