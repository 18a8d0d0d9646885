"""Minimal stand-in for the four Mesa symbols the reference imports.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  `mesa` is not installable in
the build container; the reference uses only `mesa.Model`, `mesa.Agent`,
`mesa.time.BaseScheduler` and `mesa.datacollection.DataCollector`
(reference: protonoc/simulator/model.py:26-28, entities.py:29).  Semantics
follow Mesa 0.8.x: the scheduler is an insertion-ordered container of live
agents; string model reporters are attribute lookups.
"""


class Model:
    def __init__(self, *args, **kwargs):
        self.running = True
        self.schedule = None
        self.current_id = 0

    def next_id(self):
        self.current_id += 1
        return self.current_id


class Agent:
    def __init__(self, unique_id, model):
        self.unique_id = unique_id
        self.model = model

    def step(self):
        pass
