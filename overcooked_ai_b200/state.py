"""Host-side value types of one Overcooked environment, as VIEWS of the engine's packed fields.

The reference's ``Recipe`` / ``ObjectState`` / ``SoupState`` / ``PlayerState`` / ``OvercookedState``
(src/overcooked_ai_py/mdp/overcooked_mdp.py:18-1015) are what callers of the drop-in surface hold, so the names,
attributes, equality rules, error types and the ``to_dict`` / ``from_dict`` wire keys below are the reference's.
The representation is the engine's: an object IS its 22-bit object code of include/ovc_b200.h (type, ingredient count,
ordered kinds, tick + 1), a recipe IS its row of the per-layout recipe tables, and every property decodes those bit
fields; ``layout.pack_state`` / ``unpack_state`` move the codes in and out of the int32 record without translating
them.  No game logic lives here: transitions happen on the GPU.

Deliberate difference: there is no global ``Recipe.configure`` class state (reference quirk Q1, :220-336).  Recipe
values and cook times are per-layout constants (``layout.CompiledLayout``); a soup that was built without a cook time
(legacy dicts, ``get_soup``) reports ``DEFAULT_COOK_TIME``, the reference's unconfigured ``Recipe.time`` (:163-188),
until ``unpack_state`` / ``OvercookedGridworld.soup_cook_time`` gives it its layout's.
"""
from overcooked_ai_b200.actions import Direction

ONION = "onion"
TOMATO = "tomato"
ALL_INGREDIENTS = [ONION, TOMATO]
MAX_NUM_INGREDIENTS = 3  # "num_items_for_soup" never reaches Recipe.configure (quirk Q2)
DEFAULT_COOK_TIME = 20   # Recipe.time with nothing configured (:188)

# object code fields (include/ovc_b200.h)
_TYPE = {ONION: 1, TOMATO: 2, "dish": 3, "soup": 4}
_NAME = {v: k for k, v in _TYPE.items()}
_T_SOUP = 4
_MAX_TICK = 16382


class Recipe(object):
    """An unordered multiset of 1..3 ingredients (reference :18-116), held as (n_onion, n_tomato)."""

    __slots__ = ("_o", "_t")

    def __init__(self, ingredients):
        names = tuple(ingredients)
        if not 0 < len(names) <= MAX_NUM_INGREDIENTS:
            raise ValueError("Recipe must have 1..%d ingredients" % MAX_NUM_INGREDIENTS)
        bad = [i for i in names if i not in _TYPE or _TYPE[i] > 2]
        if bad:
            raise ValueError("Invalid ingredient: %r" % (bad[0],))
        self._o, self._t = names.count(ONION), names.count(TOMATO)

    @property
    def counts(self):
        """(n_onion, n_tomato)"""
        return (self._o, self._t)

    @property
    def index(self):
        """Row of the per-layout recipe tables (cook_time / deliver_value / best_value): n_onion * 4 + n_tomato."""
        return self._o * 4 + self._t

    @property
    def ingredients(self):
        return (ONION,) * self._o + (TOMATO,) * self._t  # sorted: "onion" < "tomato"

    _ingredients = ingredients

    def __int__(self):
        # the reference's ordering key (:71-81); sorted(all_orders) depends on it
        base = MAX_NUM_INGREDIENTS + 1
        enc = self._o + base * self._t
        return (enc * base ** len(ALL_INGREDIENTS) if self._o and self._t else 0) + enc

    def __hash__(self):
        return hash(self.ingredients)

    def __eq__(self, other):
        return isinstance(other, Recipe) and self.counts == other.counts

    def __ne__(self, other):
        return not self == other

    def __lt__(self, other):
        return int(self) < int(other)

    def __repr__(self):
        return repr(self.ingredients)

    def __iter__(self):
        return iter(self.ingredients)

    def to_dict(self):
        return {"ingredients": self.ingredients}

    @classmethod
    def from_dict(cls, d):
        return cls(d["ingredients"])

    @staticmethod
    def all_recipes():
        return [Recipe([ONION] * (n - t) + [TOMATO] * t) for n in range(1, MAX_NUM_INGREDIENTS + 1) for t in range(n + 1)]


class ObjectState(object):
    """A loose or held onion / tomato / dish (reference :384-430): ``code`` is its object code (1, 2, 3)."""

    def __init__(self, name, position, **kwargs):
        self.code = _TYPE.get(name, 0)
        self._other_name = None if self.code else name  # not an engine object: kept so that is_valid() can say so
        self._position = tuple(position)

    @property
    def name(self):
        return _NAME.get(self.code & 7, self._other_name)

    @property
    def position(self):
        return self._position

    @position.setter
    def position(self, new_pos):
        self._position = tuple(new_pos)

    def is_valid(self):
        return 1 <= self.code <= 3

    def deepcopy(self):
        return ObjectState(self.name, self._position)

    def __eq__(self, other):
        return isinstance(other, ObjectState) and (self.name, self._position) == (other.name, other._position)

    def __hash__(self):
        return hash((self.name, self._position))

    def __repr__(self):
        return "{}@{}".format(self.name, self._position)

    def to_dict(self):
        return {"name": self.name, "position": self._position}

    @classmethod
    def from_dict(cls, obj_dict):
        return ObjectState(obj_dict["name"], obj_dict["position"])


class SoupState(ObjectState):
    """A soup (reference :433-693): ``code`` carries the ingredient count (bits 3-4), the ORDERED kinds (bit 5 + i set =
    slot i is a tomato; ``__eq__`` is order sensitive, quirk Q6) and ``_cooking_tick + 1`` (bits 8-21)."""

    def __init__(self, position, ingredients=None, cooking_tick=-1, cook_time=None, **kwargs):
        self._position = tuple(position)
        self._other_name = None
        items = list(ingredients) if ingredients is not None else []
        self._overfull = items[MAX_NUM_INGREDIENTS:]  # cannot be encoded; kept only so is_valid() / len() tell the truth
        kinds = 0
        for i, ing in enumerate(items[:MAX_NUM_INGREDIENTS]):
            if ing.name == TOMATO:
                kinds |= 1 << i
            elif ing.name != ONION:
                raise ValueError("invalid ingredient %r" % (ing.name,))
        self._stray = any(tuple(ing.position) != self._position for ing in items)  # is_valid (:553-560) looks at this
        self.code = _T_SOUP | (min(len(items), MAX_NUM_INGREDIENTS) << 3) | (kinds << 5)
        self._cooking_tick = cooking_tick
        self._cook_time = cook_time

    # ---- bit fields ----
    @property
    def _cooking_tick(self):
        return ((self.code >> 8) & 0x3FFF) - 1

    @_cooking_tick.setter
    def _cooking_tick(self, tick):
        if not -1 <= tick <= _MAX_TICK:
            raise ValueError("cooking tick %d outside -1..%d" % (tick, _MAX_TICK))
        self.code = (self.code & 0xFF) | ((int(tick) + 1) << 8)

    @property
    def _n(self):
        return (self.code >> 3) & 3

    @property
    def ingredients(self):
        kinds = self.code >> 5
        return [TOMATO if (kinds >> i) & 1 else ONION for i in range(self._n)] + [i.name for i in self._overfull]

    @property
    def _ingredients(self):
        return [ObjectState(name, self._position) for name in self.ingredients]

    @property
    def position(self):
        return self._position

    @position.setter
    def position(self, new_pos):  # the ingredients travel with the soup (:484-488): they are derived from it here
        self._position = tuple(new_pos)
        self._stray = False

    # ---- the reference's predicates ----
    @property
    def is_idle(self):
        return (self.code >> 8) & 0x3FFF == 0

    @property
    def cook_time(self):
        return DEFAULT_COOK_TIME if self._cook_time is None else self._cook_time

    @property
    def is_ready(self):
        return not self.is_idle and self._cooking_tick >= self.cook_time

    @property
    def is_cooking(self):
        return not self.is_idle and not self.is_ready

    @property
    def cook_time_remaining(self):
        return max(0, self.cook_time - self._cooking_tick)

    @property
    def is_full(self):
        return not self.is_idle or len(self.ingredients) == MAX_NUM_INGREDIENTS

    @property
    def recipe(self):
        if self.is_idle:
            raise ValueError("Recipe is not determined until soup begins cooking")
        return Recipe(self.ingredients)

    def is_valid(self):
        return not self._stray and not self._overfull

    def deepcopy(self):
        twin = SoupState(self._position, None, -1, self._cook_time)
        twin.code, twin._overfull, twin._stray = self.code, list(self._overfull), self._stray
        return twin

    # equality exactly as the reference defines it (:458-477): zip() truncates to the shorter ingredient list
    def __eq__(self, other):
        if not isinstance(other, SoupState) or self._position != other._position or self._cooking_tick != other._cooking_tick:
            return False
        return all(a == b for a, b in zip(self.ingredients, other.ingredients))

    def __hash__(self):
        return hash((self._position, self._cooking_tick, tuple(self.ingredients)))

    def __repr__(self):
        return "{}\nIngredients:\t{}\nCooking Tick:\t{}".format(ObjectState.__repr__(self), self._ingredients, self._cooking_tick)

    # ---- wire format (:615-663) ----
    def to_dict(self):
        idle = self.is_idle
        return {
            "name": "soup", "position": self._position,
            "_ingredients": [{"name": n, "position": self._position} for n in self.ingredients],
            "cooking_tick": self._cooking_tick, "is_cooking": self.is_cooking, "is_ready": self.is_ready, "is_idle": idle,
            "cook_time": -1 if idle else self.cook_time,
            "_cooking_tick": self._cooking_tick,  # kept for overcooked-demo, as in the reference
        }

    @classmethod
    def from_dict(cls, obj_dict):
        if obj_dict["name"] != "soup":
            return ObjectState.from_dict(obj_dict)
        if "state" in obj_dict:  # the 2019 format (:638-656): (ingredient, how many, time cooked); 20 = done
            kind, num, time = obj_dict["state"]
            # like the reference, the tomato branch leaves get_soup's default of ONE onion in place
            which = {"num_tomatoes": num} if kind == TOMATO else {"num_onions": num}
            return cls.get_soup(obj_dict["position"], cooking_tick=time or -1, finished=time >= 20, **which)
        tick = obj_dict.get("cooking_tick", obj_dict.get("_cooking_tick", -1))
        cook_time = obj_dict.get("cook_time", None)
        return cls(obj_dict["position"], [ObjectState.from_dict(i) for i in obj_dict["_ingredients"]], tick,
                   cook_time if cook_time is not None and cook_time >= 0 else None)

    @classmethod
    def get_soup(cls, position, num_onions=1, num_tomatoes=0, cooking_tick=-1, finished=False, cook_time=None, **kwargs):
        total = num_onions + num_tomatoes
        if num_onions < 0 or num_tomatoes < 0:
            raise ValueError("Number of active ingredients must be positive")
        if total > MAX_NUM_INGREDIENTS:
            raise ValueError("Too many ingredients specified for this soup")
        if cooking_tick >= 0 and total == 0:
            raise ValueError("_cooking_tick must be -1 for empty soup")
        if finished and total == 0:
            raise ValueError("Empty soup cannot be finished")
        soup = cls(position, [ObjectState(ONION, position)] * num_onions + [ObjectState(TOMATO, position)] * num_tomatoes,
                   cooking_tick, cook_time)
        if finished:  # auto_finish (:565-569): the tick jumps to the cook time
            soup._cooking_tick = soup.cook_time
        return soup


class PlayerState(object):
    """Position, facing direction and held object of one chef (reference :696-781)."""

    def __init__(self, position, orientation, held_object=None):
        self.position, self.orientation, self.held_object = tuple(position), tuple(orientation), held_object
        assert self.orientation in Direction.ALL_DIRECTIONS
        assert held_object is None or (isinstance(held_object, ObjectState) and held_object.position == self.position)

    @property
    def pos_and_or(self):
        return (self.position, self.orientation)

    def has_object(self):
        return self.held_object is not None

    def get_object(self):
        assert self.held_object is not None
        return self.held_object

    def set_object(self, obj):
        assert self.held_object is None
        obj.position = self.position
        self.held_object = obj

    def remove_object(self):
        assert self.held_object is not None
        obj, self.held_object = self.held_object, None
        return obj

    def deepcopy(self):
        return PlayerState(self.position, self.orientation, self.held_object and self.held_object.deepcopy())

    def _key(self):
        return (self.position, self.orientation, self.held_object)

    def __eq__(self, other):
        return isinstance(other, PlayerState) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return "{} facing {} holding {}".format(self.position, self.orientation, str(self.held_object))

    def to_dict(self):
        return {"position": self.position, "orientation": self.orientation,
                "held_object": self.held_object.to_dict() if self.held_object is not None else None}

    @staticmethod
    def from_dict(player_dict):
        held = player_dict.get("held_object")
        return PlayerState(player_dict["position"], player_dict["orientation"], SoupState.from_dict(held) if held is not None else None)


def _as_recipes(orders):
    return [o if isinstance(o, Recipe) else Recipe.from_dict(o) for o in orders]


class OvercookedState(object):
    """Players + loose objects + order lists + timestep (reference :784-1015)."""

    def __init__(self, players, objects, bonus_orders=[], all_orders=[], timestep=0, **kwargs):
        assert all(obj.position == pos for pos, obj in objects.items())
        self.players, self.objects, self.timestep = tuple(players), objects, timestep
        self._bonus_orders, self._all_orders = _as_recipes(bonus_orders), _as_recipes(all_orders)
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
        return tuple(p.pos_and_or for p in self.players)

    @property
    def all_orders(self):
        return sorted(self._all_orders or Recipe.all_recipes())

    @property
    def bonus_orders(self):
        return sorted(self._bonus_orders)

    def has_object(self, pos):
        return tuple(pos) in self.objects

    def get_object(self, pos):
        return self.objects[tuple(pos)]

    def add_object(self, obj, pos=None):
        pos = obj.position if pos is None else tuple(pos)
        assert pos not in self.objects
        obj.position = pos
        self.objects[pos] = obj

    def remove_object(self, pos):
        return self.objects.pop(tuple(pos))

    @classmethod
    def from_players_pos_and_or(cls, players_pos_and_or, bonus_orders=[], all_orders=[]):
        return cls([PlayerState(*po) for po in players_pos_and_or], objects={}, bonus_orders=bonus_orders, all_orders=all_orders)

    @classmethod
    def from_player_positions(cls, player_positions, bonus_orders=[], all_orders=[]):
        return cls.from_players_pos_and_or([(pos, Direction.NORTH) for pos in player_positions], bonus_orders, all_orders)

    def deepcopy(self):
        return OvercookedState([p.deepcopy() for p in self.players], {pos: obj.deepcopy() for pos, obj in self.objects.items()},
                               bonus_orders=self.bonus_orders, all_orders=self.all_orders, timestep=self.timestep)

    def time_independent_equal(self, other):
        return (isinstance(other, OvercookedState) and self.players == other.players
                and set(self.objects.items()) == set(other.objects.items())
                and (self.all_orders, self.bonus_orders) == (other.all_orders, other.bonus_orders))

    def __eq__(self, other):
        return self.time_independent_equal(other) and self.timestep == other.timestep

    def __hash__(self):
        return hash((self.players, tuple(self.objects.values()), hash(tuple(self.bonus_orders)) + hash(tuple(self.all_orders))))

    def __str__(self):
        return "Players: {}, Objects: {}, Bonus orders: {} All orders: {} Timestep: {}".format(
            str(self.players), str(list(self.objects.values())), str(self.bonus_orders), str(self.all_orders), str(self.timestep))

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
        objs = [SoupState.from_dict(o) for o in state_dict["objects"]]
        return OvercookedState([PlayerState.from_dict(p) for p in state_dict["players"]], {o.position: o for o in objs},
                               bonus_orders=state_dict.get("bonus_orders", []), all_orders=state_dict.get("all_orders", []),
                               timestep=state_dict.get("timestep", 0))
