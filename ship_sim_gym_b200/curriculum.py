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
