"""BeamProp: the 15 section scalars the line elements read (reference: pyfe3d/beamprop.pxd:1-3)."""


class BeamProp:
    FIELDS = ["A", "E", "G", "Iyy", "Izz", "Iyz", "J", "Ay", "Az",
              "intrho", "intrhoy", "intrhoz", "intrhoy2", "intrhoz2", "intrhoyz"]

    def __init__(self):
        for f in self.FIELDS:
            setattr(self, f, 0.)
