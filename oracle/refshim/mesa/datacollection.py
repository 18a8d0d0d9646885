"""DataCollector stand-in: model reporters by attribute name or callable."""
import pandas as pd


class DataCollector:
    def __init__(self, model_reporters=None, agent_reporters=None, tables=None):
        self.model_reporters = dict(model_reporters or {})
        self.agent_reporters = dict(agent_reporters or {})
        self.model_vars = {k: [] for k in self.model_reporters}
        self._agent_records = {}

    def collect(self, model):
        for name, rep in self.model_reporters.items():
            if callable(rep):
                self.model_vars[name].append(rep(model))
            else:
                self.model_vars[name].append(getattr(model, rep, None))
        if self.agent_reporters:
            recs = []
            for agent in model.schedule.agents:
                row = [model.schedule.steps, agent.unique_id]
                for name, rep in self.agent_reporters.items():
                    row.append(rep(agent) if callable(rep) else getattr(agent, rep, None))
                recs.append(tuple(row))
            self._agent_records[model.schedule.steps] = recs

    def get_model_vars_dataframe(self):
        return pd.DataFrame(self.model_vars)

    def get_agent_vars_dataframe(self):
        rows = [r for recs in self._agent_records.values() for r in recs]
        cols = ["Step", "AgentID"] + list(self.agent_reporters)
        df = pd.DataFrame.from_records(rows, columns=cols)
        return df.set_index(["Step", "AgentID"])
