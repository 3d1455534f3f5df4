"""Host-side value types of one Overcooked environment.

These mirror the reference's ``Recipe`` / ``ObjectState`` / ``SoupState`` / ``PlayerState`` /
``OvercookedState`` (src/overcooked_ai_py/mdp/overcooked_mdp.py:18-1015) closely enough that the
drop-in adapters can hand callers objects with the same attributes, equality rules and
``to_dict`` / ``from_dict`` wire format.  They carry no game logic: transitions happen on the
GPU on the packed int32 record (see include/ovc_b200.h and layout.pack_state / unpack_state).

Deliberate difference: there is no global ``Recipe.configure`` class state (reference quirk Q1,
overcooked_mdp.py:220-336).  Recipe value / cook time live in the per-layout constant table, so
several layouts can be alive in one process without changing each other's rewards.
"""
import copy

from overcooked_ai_b200.actions import Direction

ONION = "onion"
TOMATO = "tomato"
ALL_INGREDIENTS = [ONION, TOMATO]
MAX_NUM_INGREDIENTS = 3  # "num_items_for_soup" never reaches Recipe.configure (quirk Q2)


class Recipe(object):
    """An unordered multiset of 1..3 ingredients (reference :18-116)."""

    __slots__ = ("_ingredients",)

    def __init__(self, ingredients):
        ingredients = tuple(ingredients)
        if not 0 < len(ingredients) <= MAX_NUM_INGREDIENTS:
            raise ValueError("Recipe must have 1..%d ingredients" % MAX_NUM_INGREDIENTS)
        for i in ingredients:
            if i not in ALL_INGREDIENTS:
                raise ValueError("Invalid ingredient: %r" % (i,))
        self._ingredients = tuple(sorted(ingredients))

    @property
    def ingredients(self):
        return self._ingredients

    @property
    def counts(self):
        """(n_onion, n_tomato)"""
        return (self._ingredients.count(ONION), self._ingredients.count(TOMATO))

    @property
    def index(self):
        """Row of the per-layout recipe tables: n_onion * 4 + n_tomato."""
        o, t = self.counts
        return o * 4 + t

    def __int__(self):
        # same ordering key as the reference (:71-81), used by sorted(all_orders)
        o, t = self.counts
        enc = o + (MAX_NUM_INGREDIENTS + 1) * t
        return int(bool(o * t)) * enc * (MAX_NUM_INGREDIENTS + 1) ** len(ALL_INGREDIENTS) + enc

    def __hash__(self):
        return hash(self._ingredients)

    def __eq__(self, other):
        return isinstance(other, Recipe) and self._ingredients == other._ingredients

    def __ne__(self, other):
        return not self == other

    def __lt__(self, other):
        return int(self) < int(other)

    def __repr__(self):
        return repr(self._ingredients)

    def __iter__(self):
        return iter(self._ingredients)

    def to_dict(self):
        return {"ingredients": self._ingredients}

    @classmethod
    def from_dict(cls, d):
        return cls(d["ingredients"])

    @staticmethod
    def all_recipes():
        out = []
        for n in range(1, MAX_NUM_INGREDIENTS + 1):
            for t in range(0, n + 1):
                out.append(Recipe([ONION] * (n - t) + [TOMATO] * t))
        return out


class ObjectState(object):
    """A loose or held onion / tomato / dish (reference :384-430)."""

    def __init__(self, name, position, **kwargs):
        self.name = name
        self._position = tuple(position)

    @property
    def position(self):
        return self._position

    @position.setter
    def position(self, new_pos):
        self._position = tuple(new_pos)

    def is_valid(self):
        return self.name in ("onion", "tomato", "dish")

    def deepcopy(self):
        return ObjectState(self.name, self.position)

    def __eq__(self, other):
        return (
            isinstance(other, ObjectState)
            and self.name == other.name
            and self.position == other.position
        )

    def __hash__(self):
        return hash((self.name, self.position))

    def __repr__(self):
        return "{}@{}".format(self.name, self.position)

    def to_dict(self):
        return {"name": self.name, "position": self.position}

    @classmethod
    def from_dict(cls, obj_dict):
        return ObjectState(obj_dict["name"], obj_dict["position"])


class SoupState(ObjectState):
    """A soup: ordered ingredient list + cooking tick (reference :433-693).

    ``cook_time`` is a per-layout constant in this engine; ``unpack_state`` fills it in from the
    layout's recipe table, ``from_dict`` takes it from the dict like the reference does.
    """

    def __init__(self, position, ingredients=None, cooking_tick=-1, cook_time=None, **kwargs):
        super(SoupState, self).__init__("soup", position)
        self._ingredients = list(ingredients) if ingredients is not None else []
        self._cooking_tick = cooking_tick
        self._cook_time = cook_time

    # -- equality exactly as the reference defines it (:458-477): zip truncates (quirk Q6)
    def __eq__(self, other):
        return (
            isinstance(other, SoupState)
            and self.name == other.name
            and self.position == other.position
            and self._cooking_tick == other._cooking_tick
            and all(a == b for a, b in zip(self._ingredients, other._ingredients))
        )

    def __hash__(self):
        return hash(
            (
                ObjectState.__hash__(self),
                self._cooking_tick,
                hash(tuple(hash(i) for i in self._ingredients)),
            )
        )

    def __repr__(self):
        return "{}\nIngredients:\t{}\nCooking Tick:\t{}".format(
            ObjectState.__repr__(self), self._ingredients, self._cooking_tick
        )

    @property
    def position(self):
        return self._position

    @position.setter
    def position(self, new_pos):
        self._position = tuple(new_pos)
        for ing in self._ingredients:
            ing.position = new_pos

    @property
    def ingredients(self):
        return [i.name for i in self._ingredients]

    @property
    def recipe(self):
        if self.is_idle:
            raise ValueError("Recipe is not determined until soup begins cooking")
        return Recipe(self.ingredients)

    @property
    def cook_time(self):
        if self._cook_time is None:
            raise ValueError(
                "cook_time of this soup is unknown: it is a per-layout constant, "
                "use OvercookedGridworld.soup_cook_time(soup)"
            )
        return self._cook_time

    @property
    def is_idle(self):
        return self._cooking_tick < 0

    @property
    def is_ready(self):
        return (not self.is_idle) and self._cooking_tick >= self.cook_time

    @property
    def is_cooking(self):
        return not self.is_idle and not self.is_ready

    @property
    def cook_time_remaining(self):
        return max(0, self.cook_time - self._cooking_tick)

    @property
    def is_full(self):
        return not self.is_idle or len(self._ingredients) == MAX_NUM_INGREDIENTS

    def is_valid(self):
        if not all(i.position == self.position for i in self._ingredients):
            return False
        return len(self._ingredients) <= MAX_NUM_INGREDIENTS

    def deepcopy(self):
        return SoupState(
            self.position,
            [i.deepcopy() for i in self._ingredients],
            self._cooking_tick,
            self._cook_time,
        )

    def to_dict(self):
        d = ObjectState.to_dict(self)
        d["_ingredients"] = [i.to_dict() for i in self._ingredients]
        d["cooking_tick"] = self._cooking_tick
        d["is_cooking"] = self.is_cooking
        d["is_ready"] = self.is_ready
        d["is_idle"] = self.is_idle
        d["cook_time"] = -1 if self.is_idle else self.cook_time
        d["_cooking_tick"] = self._cooking_tick  # kept for overcooked-demo, as in the reference
        return d

    @classmethod
    def from_dict(cls, obj_dict):
        obj_dict = copy.deepcopy(obj_dict)
        if obj_dict["name"] != "soup":
            return ObjectState.from_dict(obj_dict)
        if "state" in obj_dict:
            # legacy (2019) soup representation, reference :638-656
            ingredient, num, time = obj_dict["state"]
            tick = -1 if time == 0 else time
            n_t = num if ingredient == TOMATO else 0
            n_o = 0 if ingredient == TOMATO else num
            return SoupState.get_soup(
                obj_dict["position"], n_o, n_t, cooking_tick=tick, finished=time >= 20
            )
        ings = [ObjectState.from_dict(i) for i in obj_dict["_ingredients"]]
        tick = obj_dict.get("cooking_tick", obj_dict.get("_cooking_tick", -1))
        cook_time = obj_dict.get("cook_time", None)
        if cook_time is not None and cook_time < 0:
            cook_time = None
        return cls(obj_dict["position"], ings, tick, cook_time)

    @classmethod
    def get_soup(
        cls, position, num_onions=1, num_tomatoes=0, cooking_tick=-1, finished=False,
        cook_time=None, **kwargs
    ):
        if num_onions < 0 or num_tomatoes < 0:
            raise ValueError("Number of active ingredients must be positive")
        if num_onions + num_tomatoes > MAX_NUM_INGREDIENTS:
            raise ValueError("Too many ingredients specified for this soup")
        if cooking_tick >= 0 and num_onions + num_tomatoes == 0:
            raise ValueError("_cooking_tick must be -1 for empty soup")
        if finished and num_onions + num_tomatoes == 0:
            raise ValueError("Empty soup cannot be finished")
        ings = [ObjectState(ONION, position) for _ in range(num_onions)]
        ings += [ObjectState(TOMATO, position) for _ in range(num_tomatoes)]
        soup = cls(position, ings, cooking_tick, cook_time)
        if finished:
            # auto_finish (:565-569): tick := cook_time; needs the layout's cook time
            soup._cooking_tick = soup.cook_time
        return soup


class PlayerState(object):
    """Position, facing direction and held object of one chef (reference :696-781)."""

    def __init__(self, position, orientation, held_object=None):
        self.position = tuple(position)
        self.orientation = tuple(orientation)
        self.held_object = held_object
        assert self.orientation in Direction.ALL_DIRECTIONS
        if self.held_object is not None:
            assert isinstance(self.held_object, ObjectState)
            assert self.held_object.position == self.position

    @property
    def pos_and_or(self):
        return (self.position, self.orientation)

    def has_object(self):
        return self.held_object is not None

    def get_object(self):
        assert self.has_object()
        return self.held_object

    def set_object(self, obj):
        assert not self.has_object()
        obj.position = self.position
        self.held_object = obj

    def remove_object(self):
        assert self.has_object()
        obj, self.held_object = self.held_object, None
        return obj

    def deepcopy(self):
        held = None if self.held_object is None else self.held_object.deepcopy()
        return PlayerState(self.position, self.orientation, held)

    def __eq__(self, other):
        return (
            isinstance(other, PlayerState)
            and self.position == other.position
            and self.orientation == other.orientation
            and self.held_object == other.held_object
        )

    def __hash__(self):
        return hash((self.position, self.orientation, self.held_object))

    def __repr__(self):
        return "{} facing {} holding {}".format(
            self.position, self.orientation, str(self.held_object)
        )

    def to_dict(self):
        return {
            "position": self.position,
            "orientation": self.orientation,
            "held_object": None if self.held_object is None else self.held_object.to_dict(),
        }

    @staticmethod
    def from_dict(player_dict):
        held = player_dict.get("held_object", None)
        if held is not None:
            held = SoupState.from_dict(held)
        return PlayerState(player_dict["position"], player_dict["orientation"], held)


class OvercookedState(object):
    """Players + loose objects + order lists + timestep (reference :784-1015)."""

    def __init__(self, players, objects, bonus_orders=[], all_orders=[], timestep=0, **kwargs):
        for pos, obj in objects.items():
            assert obj.position == pos
        self.players = tuple(players)
        self.objects = objects
        self._bonus_orders = [o if isinstance(o, Recipe) else Recipe.from_dict(o) for o in bonus_orders]
        self._all_orders = [o if isinstance(o, Recipe) else Recipe.from_dict(o) for o in all_orders]
        self.timestep = timestep
        assert len(set(self._bonus_orders)) == len(self._bonus_orders), "Bonus orders must not have duplicates"
        assert len(set(self._all_orders)) == len(self._all_orders), "All orders must not have duplicates"
        assert set(self.bonus_orders).issubset(set(self.all_orders)), "Bonus orders must be a subset of all orders"

    @property
    def player_positions(self):
        return tuple(p.position for p in self.players)

    @property
    def player_orientations(self):
        return tuple(p.orientation for p in self.players)

    @property
    def players_pos_and_or(self):
        return tuple(zip(self.player_positions, self.player_orientations))

    @property
    def all_orders(self):
        return sorted(self._all_orders) if self._all_orders else sorted(Recipe.all_recipes())

    @property
    def bonus_orders(self):
        return sorted(self._bonus_orders)

    def has_object(self, pos):
        return tuple(pos) in self.objects

    def get_object(self, pos):
        return self.objects[tuple(pos)]

    def add_object(self, obj, pos=None):
        pos = obj.position if pos is None else tuple(pos)
        assert not self.has_object(pos)
        obj.position = pos
        self.objects[pos] = obj

    def remove_object(self, pos):
        return self.objects.pop(tuple(pos))

    @classmethod
    def from_players_pos_and_or(cls, players_pos_and_or, bonus_orders=[], all_orders=[]):
        return cls(
            [PlayerState(*pos_or) for pos_or in players_pos_and_or],
            objects={},
            bonus_orders=bonus_orders,
            all_orders=all_orders,
        )

    @classmethod
    def from_player_positions(cls, player_positions, bonus_orders=[], all_orders=[]):
        return cls.from_players_pos_and_or(
            [(pos, Direction.NORTH) for pos in player_positions], bonus_orders, all_orders
        )

    def deepcopy(self):
        return OvercookedState(
            players=[p.deepcopy() for p in self.players],
            objects={pos: obj.deepcopy() for pos, obj in self.objects.items()},
            bonus_orders=[o.to_dict() for o in self.bonus_orders],
            all_orders=[o.to_dict() for o in self.all_orders],
            timestep=self.timestep,
        )

    def time_independent_equal(self, other):
        return (
            isinstance(other, OvercookedState)
            and self.players == other.players
            and set(self.objects.items()) == set(other.objects.items())
            and self.all_orders == other.all_orders
            and self.bonus_orders == other.bonus_orders
        )

    def __eq__(self, other):
        return self.time_independent_equal(other) and self.timestep == other.timestep

    def __hash__(self):
        order_hash = hash(tuple(self.bonus_orders)) + hash(tuple(self.all_orders))
        return hash((self.players, tuple(self.objects.values()), order_hash))

    def __str__(self):
        return "Players: {}, Objects: {}, Bonus orders: {} All orders: {} Timestep: {}".format(
            str(self.players),
            str(list(self.objects.values())),
            str(self.bonus_orders),
            str(self.all_orders),
            str(self.timestep),
        )

    def to_dict(self):
        return {
            "players": [p.to_dict() for p in self.players],
            "objects": [obj.to_dict() for obj in self.objects.values()],
            "bonus_orders": [o.to_dict() for o in self.bonus_orders],
            "all_orders": [o.to_dict() for o in self.all_orders],
            "timestep": self.timestep,
        }

    @staticmethod
    def from_dict(state_dict):
        players = [PlayerState.from_dict(p) for p in state_dict["players"]]
        objs = [SoupState.from_dict(o) for o in state_dict["objects"]]
        return OvercookedState(
            players,
            {o.position: o for o in objs},
            bonus_orders=state_dict.get("bonus_orders", []),
            all_orders=state_dict.get("all_orders", []),
            timestep=state_dict.get("timestep", 0),
        )
