"""oapackage: the two classes the reference's Pareto-front code touches (dataset.py:82-87, 358-365), restated as a
brute-force multi-objective (maximise every coordinate) front.  Used only to run the REFERENCE's augmentation in the
pinning tests; the product computes the front itself (osrl_b200.common.dataset.pareto_front_2d)."""


def doubleVector(values):
    return tuple(float(v) for v in values)


class ParetoDoubleLong:
    def __init__(self):
        self.items = []

    def addvalue(self, vec, index):
        self.items.append((tuple(vec), int(index)))

    def show(self, verbose=1):
        pass

    def allindices(self):
        out = []
        for v, i in self.items:
            dominated = any(all(w[k] >= v[k] for k in range(len(v))) and any(w[k] > v[k] for k in range(len(v)))
                            for w, _ in self.items)
            if not dominated:
                out.append(i)
        return tuple(sorted(out))
