"""Minimal stand-in for the `prettytable` package (absent from this image): the three things evaluate.py uses --
PrettyTable(), .field_names, .add_row(), str().  Presentation only."""


class PrettyTable(object):
    def __init__(self):
        self.field_names = []
        self._rows = []

    def add_row(self, row):
        self._rows.append([str(c) for c in row])

    def __str__(self):
        head = [str(c) for c in self.field_names]
        width = [max(len(r[i]) for r in [head] + self._rows) for i in range(len(head))]
        bar = "+" + "+".join("-" * (w + 2) for w in width) + "+"
        fmt = lambda r: "|" + "|".join(" " + c.center(w) + " " for c, w in zip(r, width)) + "|"
        return "\n".join([bar, fmt(head), bar] + [fmt(r) for r in self._rows] + [bar])
