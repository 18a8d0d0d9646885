"""BaseScheduler stand-in (Mesa 0.8.x semantics: OrderedDict keyed by unique_id)."""
from collections import OrderedDict


class BaseScheduler:
    def __init__(self, model):
        self.model = model
        self.steps = 0
        self.time = 0
        self._agents = OrderedDict()

    def add(self, agent):
        self._agents[agent.unique_id] = agent

    def remove(self, agent):
        del self._agents[agent.unique_id]

    def step(self):
        for agent in list(self._agents.values()):
            agent.step()
        self.steps += 1
        self.time += 1

    def get_agent_count(self):
        return len(self._agents)

    @property
    def agents(self):
        return list(self._agents.values())
