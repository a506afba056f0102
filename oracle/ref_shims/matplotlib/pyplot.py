class Axes:
    pass


class Figure:
    pass
