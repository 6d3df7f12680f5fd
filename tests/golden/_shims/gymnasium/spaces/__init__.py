class Space:
    pass


class Box(Space):
    pass


class Discrete(Space):
    pass
