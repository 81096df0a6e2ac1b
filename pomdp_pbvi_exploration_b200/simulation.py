"""
Policy execution with the reference's interface (src/pomdp.py:2581-3380; RewardSet / Simulation bases src/mdp.py:1528-1978):
`RewardSet`, `SimulationHistory`, `Simulation`, `SimulationSet`, `Agent`.

The two array operations of a simulation step are the hot path's own kernels:
    best action   argmax_v b.alpha_v  ->  pbvi_max_values          (reference Agent.get_best_action, :2893-2945)
    belief update b' ~ RTO[., a, o, .] b  ->  pbvi_belief_update   (reference bincount2D_vectorized, :3277-3310)
for all running simulations at once.  States, observations and rewards are sampled on the host with NumPy's global RNG
in the reference's own draw order (one `np.random.random(n)` per step for the observations, one `np.random.choice` per
simulation for stochastic transitions), so a seeded run follows the reference's CPU trajectory.  Plots and videos are
out of scope.
"""
from __future__ import annotations

import os
from datetime import datetime
from typing import Tuple, Union

import numpy as np
import torch

from .belief import Belief
from .model import Model
from .value_function import ValueFunction


class RewardSet(list):
    """List of rewards of one simulation (reference src/mdp.py:1528-1567)."""

    def __init__(self, items: list = []):
        self.extend(items)

    def get_total_discounted_reward(self, gamma: float) -> float:
        rewards = np.array(self)
        return np.dot(rewards, gamma ** np.arange(len(self)))


class SimulationHistory:
    """States, actions, observations, rewards and (lazily re-derived) beliefs of one simulation
    (reference src/mdp.py:1689-1757 + src/pomdp.py:2581-2660)."""

    def __init__(self, model: Model, start_state: int, start_belief: Belief):
        self.model = model
        self.states = [start_state]
        self.actions = []
        self.rewards = RewardSet()
        self._beliefs = [start_belief]
        self.observations = []

    @property
    def grid_point_sequence(self) -> list:
        return [[int(i[0]) for i in np.where(self.model.state_grid == s)] for s in self.states]

    @property
    def beliefs(self) -> list:
        if len(self._beliefs) < len(self):
            chain = self.model.device.belief_trajectory(self._beliefs[0].values, [int(a) for a in self.actions],
                                                        [int(o) for o in self.observations])
            self._beliefs = [self._beliefs[0]] + [Belief._from_device(self.model, row) for row in chain]
        return self._beliefs

    def add(self, action: int, reward, next_state: int, next_belief: Belief, observation: int) -> None:
        self.actions.append(action)
        self.rewards.append(reward)
        self.states.append(next_state)
        self._beliefs.append(next_belief)
        self.observations.append(observation)

    def __len__(self):
        return len(self.states)

    def to_dataframe(self, include_beliefs: bool = False):
        import pandas as pd
        points = self.grid_point_sequence
        df = pd.DataFrame({'States': self.states,
                           'State_grid_x': [p[0] for p in points],
                           'State_grid_y': [p[1] for p in points],
                           'Actions': self.actions + [None],
                           'Rewards': list(self.rewards) + [None]})
        df['Observations'] = self.observations + [None]
        if include_beliefs:
            belief_array = np.array([b.values_host for b in self.beliefs])
            df = pd.concat([df, pd.DataFrame(belief_array, columns=[f'B_{sl}' for sl in self.model.state_labels])], axis=1)
        return df

    def save(self, path: str = './Simulations', file_name: Union[str, None] = None, include_beliefs: bool = False) -> None:
        if not os.path.exists(path):
            os.makedirs(path)
        if file_name is None:
            file_name = datetime.now().strftime('%Y%m%d_%H%M%S') + '_simulation.csv'
        if not file_name.endswith('.csv'):
            file_name += '.csv'
        self.to_dataframe(include_beliefs=include_beliefs).to_csv(path + '/' + file_name, index=False)


class Simulation:
    """One simulated agent (reference src/mdp.py:1888-1978 + src/pomdp.py:2763-2810)."""

    def __init__(self, model: Model) -> None:
        self.model = model
        self.agent_state = -1
        self.is_done = True
        self.initialize_simulation()

    def initialize_simulation(self, start_state: Union[int, None] = None) -> int:
        if start_state is None:
            self.agent_state = int(np.random.choice(a=self.model.states, size=1, p=self.model.start_probabilities)[0])
        else:
            self.agent_state = start_state
        self.is_done = False
        return self.agent_state

    def run_action(self, a: int) -> Tuple[Union[int, float], int]:
        assert not self.is_done, "Action run when simulation is done."
        s = self.agent_state
        s_p = self.model.transition(s, a)
        o = self.model.observe(s_p, a)
        r = self.model.reward(s, a, s_p, o)
        self.agent_state = s_p
        if s_p in self.model.end_states:
            self.is_done = True
        if a in self.model.end_actions:
            self.is_done = True
        return (r, o)


class SimulationSet:
    """n simulated agents stepped together (reference src/pomdp.py:2813-2950)."""

    def __init__(self, model: Model):
        self.model = model
        self.n = -1
        self.agent_states = [-1]
        self.simulations = []
        self.is_done = [True]

    def initialize_simulations(self, n: int = 1, start_state: Union[list, int, None] = None) -> np.ndarray:
        if isinstance(start_state, int):
            start_state_array = (np.ones(n) * start_state).astype(int)
        elif isinstance(start_state, list):
            repeated_list = np.repeat(np.array(start_state), int(np.ceil(n / len(start_state))))
            start_state_array = np.resize(repeated_list, n)
        else:
            start_state_array = np.random.choice(self.model.states, size=n, p=self.model.start_probabilities).astype(int)
        self.n = n
        self.agent_states = start_state_array
        self.simulations = np.arange(n)
        self.is_done = np.zeros(n, dtype=bool)
        return self.agent_states

    def _rewards(self, s, a, s_p, o) -> np.ndarray:
        m = self.model
        if m.immediate_reward_function is not None:
            return np.asarray(m.immediate_reward_function(s, a, s_p, o), dtype=float) * np.ones(len(s))
        return m.immediate_reward_table[s, a, s_p, o].astype(float)

    def run_actions(self, actions: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        m = self.model
        actions = np.asarray(actions).astype(int)
        next_state_potentials = m.reachable_states[self.agent_states, actions, :]
        if m.reachable_state_count == 1:
            next_states = next_state_potentials[:, 0]
        else:
            potential_probabilities = m.reachable_probabilities[self.agent_states, actions, :]
            chosen = np.array([np.random.choice(len(p), size=1, p=p)[0] for p in potential_probabilities], dtype=int)
            next_states = next_state_potentials[np.arange(self.n), chosen]
        observation_probabilities = m.observation_table[next_states, actions, :]
        observations = np.sum(np.random.random(self.n)[:, None] > np.cumsum(observation_probabilities[:, :-1], axis=1), axis=1)
        step_rewards = self._rewards(self.agent_states, actions, next_states, observations)
        rewards = np.where(~self.is_done, step_rewards, 0)
        self.is_done = self.is_done | np.isin(next_states, np.array(m.end_states))
        self.agent_states = next_states
        return rewards, observations


class Agent:
    """An agent acting on a value function (reference src/pomdp.py:2953-3380)."""

    def __init__(self, model: Model, value_function: Union[ValueFunction, None] = None) -> None:
        self.model = model
        self.value_function = value_function

    def train(self, solver, expansions: int, horizon: int):
        """As in the reference (:2978-3002), `horizon` lands on the third positional parameter of `solve` (full_backup)."""
        self.value_function, solve_history = solver.solve(self.model, expansions, horizon)
        return solve_history

    def get_best_action(self, belief):
        """Action of argmax_v b.alpha_v for one Belief (-> int) or a [n,S] array of beliefs (-> array), :2893-2945."""
        assert self.value_function is not None, "No value function, training probably has to be run..."
        is_single = isinstance(belief, Belief)
        rows = belief.values[None, :] if is_single else belief
        _, best = self.model.device.max_values(rows, self.value_function.alpha_vector_array)
        best_actions = self.value_function.actions[best.cpu().numpy()]
        return int(best_actions[0]) if is_single else best_actions

    def simulate(self, simulator: Union[Simulation, None] = None, max_steps: int = 1000, start_state: Union[int, None] = None,
                 initial_belief: Union[Belief, None] = None, print_progress: bool = True, print_stats: bool = True) -> SimulationHistory:
        assert self.value_function is not None, "No value function, training probably has to be run..."
        if simulator is None:
            simulator = Simulation(self.model)
        s = simulator.initialize_simulation(start_state=start_state)
        belief = Belief(self.model) if initial_belief is None else initial_belief
        history = SimulationHistory(self.model, start_state=s, start_belief=belief)
        sim_start_ts = datetime.now()
        for _ in range(max_steps):
            a = self.get_best_action(belief)
            r, o = simulator.run_action(a)
            new_belief = belief.update(a, o)
            history.add(action=a, next_state=simulator.agent_state, next_belief=new_belief, reward=r, observation=o)
            belief = new_belief
            if simulator.is_done:
                break
        if print_stats:
            print('Simulation done:')
            print(f'\t- Runtime (s): {(datetime.now() - sim_start_ts).total_seconds()}')
            print(f'\t- Steps: {len(history.states)}')
            print(f'\t- Total rewards: {sum(history.rewards)}')
            print(f'\t- End state: {self.model.state_labels[history.states[-1]]}')
        return history

    def run_n_simulations(self, simulator: Union[Simulation, None] = None, n: int = 1000, max_steps: int = 1000,
                          start_states: Union[list, int, None] = None, initial_beliefs=None, reward_discount: float = 0.99,
                          print_progress: bool = True, print_stats: bool = True):
        if simulator is None:
            simulator = Simulation(self.model)
        assert (not isinstance(start_states, list)) or (len(start_states) == n), 'The size of the list of start states has to match n'
        assert (not isinstance(initial_beliefs, list)) or (len(initial_beliefs) == n), 'The size of the list of initial beliefs has to match n'
        sim_start_ts = datetime.now()
        all_histories, all_final_rewards, all_discounted, all_len, done_count = [], RewardSet(), [], [], 0
        for i in range(n):
            hist = self.simulate(simulator=simulator, max_steps=max_steps,
                                 start_state=(start_states if not isinstance(start_states, list) else start_states[i]),
                                 initial_belief=(initial_beliefs if not isinstance(initial_beliefs, list) else initial_beliefs[i]),
                                 print_progress=False, print_stats=False)
            done_count += int(simulator.is_done)
            all_histories.append(hist)
            all_final_rewards.append(np.sum(hist.rewards))
            all_discounted.append(hist.rewards.get_total_discounted_reward(reward_discount))
            all_len.append(len(hist))
        if print_stats:
            print(f'All {n} simulations done:')
            print(f'\t- Average runtime (s): {((datetime.now() - sim_start_ts).total_seconds() / n)}')
            print(f'\t- Simulations reached goal: {done_count}/{n} ({n - done_count} failures)')
            print(f'\t- Average step count: {(sum(all_len) / n)}')
            print(f'\t- Average total rewards: {(sum(all_final_rewards) / n)}')
            print(f'\t- Average discounted rewards (ADR): {(sum(all_discounted) / n)}')
        return all_final_rewards, all_histories

    def run_n_simulations_parallel(self, n: int = 1000, simulator_set: Union[SimulationSet, None] = None, max_steps: int = 1000,
                                   start_states: Union[list, int, None] = None, initial_beliefs=None, reward_discount: float = 0.99,
                                   print_progress: bool = True, print_stats: bool = True):
        """
        n simulations stepped together (reference :3203-3380): per step one batched argmax(B.V^T), one host sampling of
        (s', o, r) for all agents, one batched belief update; finished simulations are filtered out of the batch.
        """
        model, dev = self.model, self.model.device
        assert (not isinstance(start_states, list)) or (len(start_states) == n), 'The size of the list of start states has to match n'
        assert (not isinstance(initial_beliefs, list)) or (len(initial_beliefs) == n), 'The size of the list of initial beliefs has to match n'
        if initial_beliefs is None:
            beliefs = Belief(model).values[None, :].repeat(n, 1)
        elif isinstance(initial_beliefs, Belief):
            beliefs = initial_beliefs.values[None, :].repeat(n, 1)
        else:
            beliefs = torch.stack([b.values for b in initial_beliefs])
        if simulator_set is None:
            simulator_set = SimulationSet(model)
        start_state_array = simulator_set.initialize_simulations(n, start_states)
        done_at_step = np.full(n, -1, dtype=int)
        simulations = np.arange(n)
        discount = reward_discount
        rewards_history = np.zeros((max_steps, n))
        discounted_rewards_history = np.zeros((max_steps, n))
        states_history = np.empty((max_steps + 1, n))
        states_history[0] = start_state_array
        actions_history = np.empty((max_steps, n))
        observations_history = np.empty((max_steps, n))
        sim_start_ts = datetime.now()
        for i in range(max_steps):
            best_actions = self.get_best_action(beliefs)
            rewards, observations = simulator_set.run_actions(best_actions)
            beliefs, _ = dev.belief_update(beliefs, best_actions.astype(np.int32), observations.astype(np.int32))
            rewards_history[i, simulations] = rewards
            discounted_rewards_history[i, simulations] = rewards * discount
            states_history[i + 1, simulations] = simulator_set.agent_states
            actions_history[i, simulations] = best_actions
            observations_history[i, simulations] = observations
            done = simulator_set.is_done
            done_at_step[simulations[done]] = i
            simulations = simulations[~done]
            if done.any():
                beliefs = beliefs[torch.as_tensor(np.flatnonzero(~done), device=beliefs.device)]
            simulator_set.n = len(simulations)
            simulator_set.agent_states = simulator_set.agent_states[~done]
            simulator_set.simulations = simulator_set.simulations[~done]
            simulator_set.is_done = done[~done]
            discount *= reward_discount
            if len(simulations) == 0:
                break
        sim_hist_list = []
        b0 = Belief(model)
        done_at_step_sum = 0
        for i, s0 in enumerate(start_state_array):
            sim_hist = SimulationHistory(model, int(s0), b0)
            last_step = int(done_at_step[i]) if done_at_step[i] >= 0 else max_steps
            done_at_step_sum += last_step
            sim_hist.states = states_history[:last_step + 1, i].tolist()
            sim_hist.actions = actions_history[:last_step, i].tolist()
            sim_hist.observations = observations_history[:last_step, i].tolist()
            sim_hist.rewards = rewards_history[:last_step, i].tolist()
            sim_hist_list.append(sim_hist)
        done_sim_count = int(np.sum(done_at_step >= 0))
        if print_stats:
            print(f'All {n} simulations done in {(datetime.now() - sim_start_ts).total_seconds():.3f}s:')
            print(f'\t- Simulations reached goal: {done_sim_count}/{n} ({n - done_sim_count} failures)')
            print(f'\t- Average step count: {(done_at_step_sum / n)}')
            print(f'\t- Average total rewards: {(np.sum(rewards_history) / n)}')
            print(f'\t- Average discounted rewards (ADR): {(np.sum(discounted_rewards_history) / n)}')
        return RewardSet(np.sum(rewards_history, axis=0).tolist()), sim_hist_list
