"""Stub of pymunk.pygame_util (imported at game.py:11, used only for debug drawing at game.py:202-204).
TEST INFRASTRUCTURE ONLY."""


class DrawOptions(object):
    def __init__(self, surface=None):
        self.surface = surface
        self.flags = 0
