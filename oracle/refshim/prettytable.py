"""PrettyTable stand-in (only used by the reference's ProtonOC.overview)."""


class PrettyTable:
    def __init__(self):
        self.field_names = []
        self._rows = []

    def add_row(self, row):
        self._rows.append(list(row))

    def __str__(self):
        lines = [" | ".join(str(x) for x in self.field_names)]
        lines += [" | ".join(str(x) for x in r) for r in self._rows]
        return "\n".join(lines)
