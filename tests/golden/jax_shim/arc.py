"""Stand-in for the ARC (Alkali Rydberg Calculator) package: imported at TQ:7, unused by PropagatorVL."""


class Rubidium87:  # noqa: D101
    pass
