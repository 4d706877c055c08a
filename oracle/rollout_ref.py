"""TEST INFRASTRUCTURE — line-by-line restatement of the reference's rollout collector for ONE env
(RL/ppo/game_manager.py:34-150), driven by recorded tick events instead of live policies.

PINNED: ``tests/test_rollout_vs_reference.py`` runs the reference's unchanged manager (oracle/manager_harness.py) and checks
this restatement list for list over consecutive rollouts with games ending inside the window.

The reference loop body (:78-136) is kept statement for statement; `tick()` is one iteration of the `while`
loop for this env, `reset()` is GamesAndPoliciesManager.reset (:34-56), `after_rollouts()` is :142-150.
"""
from __future__ import annotations


class RefCollector:
    def __init__(self, num_steps: int, active_player_id: int):
        self.num_steps = num_steps
        self.active = active_player_id
        self.observations, self.action_masks, self.actions, self.action_log_probs = [], [], [], []
        self.rewards, self.terminal_masks = [], []

    def reset(self, players_go: int, obs):                       # game_manager.py:34-56
        self.observations, self.action_masks, self.actions, self.action_log_probs = [], [], [], []
        self.rewards, self.terminal_masks = [], []
        self.terminal_masks.append(1.0)
        if players_go == self.active:
            self.observations.append(obs)
        self.begin_gather()

    def after_rollouts(self):                                    # game_manager.py:142-150
        self.observations = [self.observations[-1]]
        self.terminal_masks = [self.terminal_masks[-1]]
        self.actions, self.action_masks, self.action_log_probs, self.rewards = [], [], [], []
        self.begin_gather()

    def begin_gather(self):                                      # game_manager.py:76-77 (locals of every gather_rollouts call)
        self._rewards = {p: 0 for p in (1, 2, 3, 4)}
        self._done_since_prev_turn = False

    def collecting(self) -> bool:                                # game_manager.py:78
        return len(self.observations) < self.num_steps + 1

    def tick(self, players_go, action_masks, actions, action_log_probs, reward, done, n_players_go_before_reset,
             n_players_go_after_reset, obs_after, obs_after_reset):
        rewards = self._rewards
        for player_id in (1, 2, 3, 4):                           # :94-95
            rewards[player_id] += reward[player_id - 1]
        n_players_go = n_players_go_before_reset                 # :99
        obs = obs_after
        reward_updated = False
        if players_go == self.active:                            # :102-105
            self.actions.append(actions)
            self.action_log_probs.append(action_log_probs)
            self.action_masks.append(action_masks)
        if n_players_go == self.active and len(self.actions) > 0:  # :106-110
            if self._done_since_prev_turn is False:
                self.rewards.append(rewards[self.active])
                rewards[self.active] = 0.0
                reward_updated = True
        if done:                                                 # :112-124
            obs = obs_after_reset
            self.terminal_masks.append(1.0 - done)
            self._done_since_prev_turn = False
            n_players_go = n_players_go_after_reset
            if reward_updated is False:
                self.rewards.append(rewards[self.active])
            for player_id in (1, 2, 3, 4):
                rewards[player_id] = 0.0
        if n_players_go == self.active:                          # :128-133
            if done is False and self._done_since_prev_turn is False:
                self.terminal_masks.append(1.0 - done)
            self._done_since_prev_turn = False
            self.observations.append(obs)
        else:                                                    # :134-136
            if done:
                self._done_since_prev_turn = True
