"""Scalar lesson schedule with the semantics of the reference helper (ship_gym/curriculum.py:23-50), which
nothing in the reference imports.  Here `progress()` is meant to be fed the all-reduced mean episode return
(BatchedShipEnv.stats()) and `int(c)` / `float(c)` to select e.g. MAX_STEPS or a scenario-bank tier.

Kept quirks: the pass test is strict (`val > condition`), the counter is NOT reset
by a failing call, so a lesson advances on the (repeat_condition + 1)-th passing call, consecutive or not.
"""
import enum


class LessonCondition(enum.Enum):
    STEPS = 0
    REWARD = 1


class Lesson(object):
    """A bundle of thresholds; passed when every tracked value reaches its threshold (curriculum.py:7-20)."""

    def __init__(self, param_dict):
        self.param_dict = dict(param_dict)

    def pass_lesson(self, val_dict):
        return all(val_dict[k] >= v for k, v in self.param_dict.items())


class Curriculum(object):
    def __init__(self, values, conditions, repeat_condition=1):
        self.values = list(values)
        self.conditions = list(conditions)
        self.repeat_condition = repeat_condition
        self.lesson = 0
        self.repeat_reached = 0

    def __int__(self):
        return int(self.values[self.lesson])

    def __float__(self):
        return float(self.values[self.lesson])

    @property
    def value(self):
        return self.values[self.lesson]

    def progress(self, val):
        """Feed one measurement; returns True when it moved the curriculum to the next lesson."""
        if self.lesson >= len(self.conditions) or not (val > self.conditions[self.lesson]):
            return False
        self.repeat_reached += 1
        if self.repeat_reached <= self.repeat_condition:
            return False
        self.lesson += 1
        self.repeat_reached = 0
        return True


class CurriculumDriver(object):
    """Wires a `Curriculum` to a batched env (the reference defines the helper but never uses it).

    After every rollout call `update()`: the episode statistics of all ranks are summed (one all-reduce of 16 doubles),
    the global mean episode return is fed to `Curriculum.progress`, and when that advances the lesson the new lesson's
    value is applied:
      * `knob="max_steps"`: the value is the new EnvConfig.MAX_STEPS (config.py:16);
      * `knob="bank"`: the value indexes `banks`, a list of ScenarioBank difficulty tiers, which is uploaded.
    Lessons only advance when at least `min_episodes` episodes finished since the last update (otherwise the mean is
    noise and the statistics keep accumulating).
    """

    def __init__(self, env, curriculum, knob="max_steps", banks=None, min_episodes=1, all_reduce=None):
        if knob not in ("max_steps", "bank"):
            raise ValueError("knob must be 'max_steps' or 'bank'")
        if knob == "bank" and not banks:
            raise ValueError("knob='bank' needs the list of scenario-bank tiers")
        self.env, self.curriculum, self.knob, self.banks = env, curriculum, knob, banks
        self.min_episodes = min_episodes
        self.all_reduce = all_reduce
        self.history = []
        self._apply()

    def _apply(self):
        if self.knob == "max_steps":
            self.env.set_max_steps(int(self.curriculum))
        else:
            self.env.load_scenarios(self.banks[int(self.curriculum)])

    def update(self):
        """Returns (advanced, mean_return or None)."""
        stats = self.env.stats_tensor(clear=False)
        if self.all_reduce is not None:
            stats = self.all_reduce(stats.clone())
        v = stats.tolist()
        episodes, return_sum = float(v[0]), float(v[1])
        if episodes < self.min_episodes:
            return False, None
        self.env.stats_tensor(clear=True)
        mean_return = return_sum / episodes
        advanced = self.curriculum.progress(mean_return)
        self.history.append((self.curriculum.lesson, mean_return, episodes))
        if advanced:
            self._apply()
        return advanced, mean_return
