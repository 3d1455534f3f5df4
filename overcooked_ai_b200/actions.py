"""Action / Direction vocabulary of the reference, as index tables.

Mirrors the public names of the reference's ``Direction`` and ``Action`` classes
(src/overcooked_ai_py/mdp/actions.py:7-57) so joint actions written for the reference —
direction tuples, ``(0, 0)`` for stay, the string ``"interact"`` — are accepted unchanged by the
drop-in adapters.  The engine itself only sees the indices 0..5 (N, S, E, W, STAY, INTERACT).
"""


class Direction(object):
    NORTH = (0, -1)
    SOUTH = (0, 1)
    EAST = (1, 0)
    WEST = (-1, 0)
    ALL_DIRECTIONS = INDEX_TO_DIRECTION = [NORTH, SOUTH, EAST, WEST]
    DIRECTION_TO_INDEX = {d: i for i, d in enumerate(INDEX_TO_DIRECTION)}
    OPPOSITE_DIRECTIONS = {NORTH: SOUTH, SOUTH: NORTH, EAST: WEST, WEST: EAST}
    DIRECTION_TO_NAME = {NORTH: "NORTH", SOUTH: "SOUTH", EAST: "EAST", WEST: "WEST"}


class Action(object):
    STAY = (0, 0)
    INTERACT = "interact"
    ALL_ACTIONS = INDEX_TO_ACTION = Direction.INDEX_TO_DIRECTION + [STAY, INTERACT]
    ACTION_TO_INDEX = {a: i for i, a in enumerate(INDEX_TO_ACTION)}
    MOTION_ACTIONS = Direction.ALL_DIRECTIONS + [STAY]
    NUM_ACTIONS = len(ALL_ACTIONS)
    ACTION_TO_CHAR = {
        Direction.NORTH: "↑",
        Direction.SOUTH: "↓",
        Direction.EAST: "→",
        Direction.WEST: "←",
        STAY: "stay",
        INTERACT: INTERACT,
    }

    @staticmethod
    def move_in_direction(point, direction):
        return (point[0] + direction[0], point[1] + direction[1])

    @staticmethod
    def to_index(action):
        """Index of a reference-style action; raises ValueError like the reference's legality
        check (overcooked_mdp.py:1394-1398) for anything that is not one of the six actions."""
        if isinstance(action, list):
            action = tuple(action)
        try:
            return Action.ACTION_TO_INDEX[action]
        except (KeyError, TypeError):
            raise ValueError("Illegal action %r" % (action,))
