"""Procedural layouts for variable-MDP training — the reference's ``LayoutGenerator`` /
``MDPParamsGenerator`` (src/overcooked_ai_py/mdp/layout_generator.py:66-397) on the host.

A generated MDP is: dig floor cells out of an all-counter inner grid until enough of it is empty AND the
floor is one connected region (:341-357), put one of every feature type on counters that touch the floor
and then random extra features (:375-397), embed the inner grid at a random offset in a counter-filled
outer grid (:322-339), draw two distinct start cells (:399-405).  Everything random comes from numpy's
GLOBAL generator through the same calls in the same order as the reference makes them, so after
``np.random.seed(k)`` both produce the same grids (tests/golden/layout_generator.npz pins that, and a
live differential test runs where the reference tree is present).  One exception: ``generate_all_orders`` /
``generate_bonus_orders`` choose from ``Recipe.ALL_RECIPES``, a Python ``set`` whose iteration order
changes with the interpreter's string-hash seed — the reference is not reproducible there; here the
candidates are listed in the canonical order of ``Recipe.all_recipes()`` and drawn with the same
``np.random.choice`` call.

The batched engine consumes generated layouts as a POOL: ``generate_layout_pool(n, params, outer_shape)``
-> list of CompiledLayout for ``BatchedOvercookedEnv(pool, ..., random_layout=True)``, whose (auto-)reset
redraws each environment's layout from the pool on the device (ovc_random_start_t.random_layout).
The single-environment drop-in keeps the reference surface: ``LayoutGenerator.mdp_gen_fn_from_dict``
returns the ``mdp_generator_fn`` that ``OvercookedEnv(mdp_generator_fn, ...)`` calls at every reset.
"""
import copy
import random

import numpy as np

from overcooked_ai_b200.state import Recipe

EMPTY, COUNTER, ONION_DISPENSER, TOMATO_DISPENSER, POT, DISH_DISPENSER, SERVING_LOC = " ", "X", "O", "T", "P", "D", "S"

DEFAULT_FEATURE_TYPES = (POT, ONION_DISPENSER, DISH_DISPENSER, SERVING_LOC)  # tomato dispensers are opt-in (:96-101)

DEFAULT_MDP_GEN_PARAMS = {
    "inner_shape": (5, 4),
    "prop_empty": 0.95,
    "prop_feats": 0.1,
    "start_all_orders": [{"ingredients": ["onion", "onion", "onion"]}],
    "recipe_values": [20],
    "recipe_times": [20],
    "display": False,
}


def DEFAILT_PARAMS_SCHEDULE_FN(outside_information):  # (sic) the reference's spelling, :52-63
    return copy.deepcopy(DEFAULT_MDP_GEN_PARAMS)


def mdp_fn_random_choice(mdp_fn_choices):
    """:30-32"""
    assert type(mdp_fn_choices) is list and len(mdp_fn_choices) > 0
    return random.choice(mdp_fn_choices)


class MDPParamsGenerator(object):
    """:66-93 — a schedule ``outside_information -> mdp_params``."""

    def __init__(self, params_schedule_fn):
        assert callable(params_schedule_fn), "params scheduling function must be a callable"
        self.params_schedule_fn = params_schedule_fn

    @staticmethod
    def from_fixed_param(mdp_params_always):
        return MDPParamsGenerator(lambda _ignored: mdp_params_always)

    def generate(self, outside_information={}):
        assert type(outside_information) is dict
        return self.params_schedule_fn(outside_information)


# ---- grid primitives: a grid is a numpy array of one-character strings indexed [x][y] -----------------
def _interior_cell(shape):
    """get_random_interior_location :551-554: x first, then y."""
    x = np.random.randint(low=1, high=shape[0] - 1)
    y = np.random.randint(low=1, high=shape[1] - 1)
    return int(x), int(y)


def _random_floor_cell(grid):
    """get_random_empty_location :556-561: rejection sampling over interior cells."""
    while True:
        x, y = _interior_cell(grid.shape)
        if grid[x, y] == EMPTY:
            return (x, y)


def _neighbours(shape, x, y):
    for dx, dy in ((0, -1), (0, 1), (1, 0), (-1, 0)):
        nx, ny = x + dx, y + dy
        if 0 <= nx < shape[0] and 0 <= ny < shape[1]:
            yield nx, ny


def dig_connected_floor(shape, prop_empty):
    """dig_space_with_disjoint_sets :341-357: dig random interior cells until more than ``prop_empty`` of the
    interior is floor and all floor cells are 4-connected (tracked with a union-find over the dug cells)."""
    W, H = int(shape[0]), int(shape[1])
    grid = np.full((W, H), COUNTER)
    n_interior = W * H - 2 * (W + H) + 4
    parent = {}
    n_sets = n_floor = 0

    def root(c):
        while parent[c] != c:
            parent[c] = parent[parent[c]]
            c = parent[c]
        return c

    while not (float(n_floor) / n_interior > prop_empty and n_sets == 1):
        if n_floor == n_interior:
            raise ValueError("prop_empty=%r cannot be exceeded on a %dx%d grid" % (prop_empty, W, H))
        while True:
            x, y = _interior_cell((W, H))
            if grid[x, y] != EMPTY:
                break
        grid[x, y] = EMPTY
        n_floor += 1
        parent[(x, y)] = (x, y)
        n_sets += 1
        for n in _neighbours((W, H), x, y):
            if n in parent:
                a, b = root(n), root((x, y))
                if a != b:
                    parent[a] = b
                    n_sets -= 1
    return grid


def feature_cells(grid):
    """valid_feature_locations :515-523: counters with a floor cell next to them, x-major order, as an (n, 2) array."""
    cells = [(x, y) for x in range(grid.shape[0]) for y in range(grid.shape[1])
             if grid[x, y] == COUNTER and any(grid[n] == EMPTY for n in _neighbours(grid.shape, x, y))]
    return np.array(cells)


def place_features(grid, prop_features=0, feature_types=DEFAULT_FEATURE_TYPES):
    """add_features :375-397: shuffle the candidate cells, give the first len(feature_types) of them one feature
    each in order, then keep adding uniformly drawn features while placed / candidates < prop_features."""
    cells = feature_cells(grid)
    np.random.shuffle(cells)
    assert len(cells) > len(feature_types)
    placed = 0
    for x, y in cells:
        if placed < len(feature_types):
            grid[x, y] = feature_types[placed]
        elif placed / len(cells) >= prop_features:
            break
        else:
            grid[x, y] = np.random.choice(feature_types)
        placed += 1


def embed(grid, outer_shape):
    """embed_grid :322-339: copy into a counter-filled outer grid at a random offset (randint's upper bound is
    exclusive, so the inner grid never touches the far edges when there is leeway — as in the reference)."""
    outer_shape = (int(outer_shape[0]), int(outer_shape[1]))
    assert grid.shape[0] <= outer_shape[0] and grid.shape[1] <= outer_shape[1]
    out = np.full(outer_shape, COUNTER)
    x_leeway, y_leeway = outer_shape[0] - grid.shape[0], outer_shape[1] - grid.shape[1]
    sx = np.random.randint(0, x_leeway) if x_leeway else 0
    sy = np.random.randint(0, y_leeway) if y_leeway else 0
    out[sx:sx + grid.shape[0], sy:sy + grid.shape[1]] = grid
    return out


def random_start_cells(grid):
    """get_random_starting_positions :399-405"""
    pos0 = _random_floor_cell(grid)
    pos1 = _random_floor_cell(grid)
    while pos0 == pos1:
        pos0 = _random_floor_cell(grid)
    return pos0, pos1


def to_layout_rows(grid, start_positions):
    """padded_grid_to_layout_grid :306-320: rows of characters [y][x] with '1' / '2' on the start cells."""
    rows = [[str(grid[x, y]) for x in range(grid.shape[0])] for y in range(grid.shape[1])]
    for i, (x, y) in enumerate(start_positions):
        rows[y][x] = str(i + 1)
    return rows


def generate_random_recipes(n=1, min_size=2, max_size=3, ingredients=None, recipes=None, unique=True):
    """Recipe.generate_random_recipes (overcooked_mdp.py:339-377) over a canonically ordered candidate list."""
    recipes = list(Recipe.all_recipes()) if recipes is None else list(recipes)
    ingredients = set(ingredients or ("onion", "tomato"))
    assert 1 <= min_size <= max_size <= 3
    assert all(i in ("onion", "tomato") for i in ingredients)
    relevant = [r for r in recipes if min_size <= len(r.ingredients) <= max_size and all(i in ingredients for i in r.ingredients)]
    assert (not unique) or n <= len(relevant)
    picks = np.random.choice(len(relevant), n, replace=not unique)
    return [relevant[int(i)] for i in picks]


class LayoutGenerator(object):
    """:104-405.  ``generate_padded_mdp`` returns an ``overcooked_ai_b200.mdp.OvercookedGridworld``."""

    def __init__(self, mdp_params_generator, outer_shape=(5, 4)):
        self.mdp_params_generator = mdp_params_generator
        self.outer_shape = outer_shape

    @staticmethod
    def mdp_gen_fn_from_dict(mdp_params, outer_shape=None, mdp_params_schedule_fn=None):
        """:115-142 — the ``mdp_generator_fn`` for OvercookedEnv: a fixed bundled layout when ``outer_shape`` is
        None, else a generator of padded random layouts."""
        from overcooked_ai_b200.mdp import OvercookedGridworld

        if outer_shape is None:
            assert type(mdp_params) is dict and "layout_name" in mdp_params
            mdp = OvercookedGridworld.from_layout_name(**mdp_params)
            return lambda _ignored: mdp
        if mdp_params_schedule_fn is None:
            assert mdp_params is not None
            mdp_pg = MDPParamsGenerator.from_fixed_param(mdp_params_always=mdp_params)
        else:
            assert mdp_params is None, (
                "please remove the mdp_params from the variable, because mdp_params_schedule_fn exist and we will "
                "always use the schedule_fn if it exist")
            mdp_pg = MDPParamsGenerator(params_schedule_fn=mdp_params_schedule_fn)
        return LayoutGenerator(mdp_pg, outer_shape).generate_padded_mdp

    def generate_padded_mdp(self, outside_information={}):
        """:144-195"""
        from overcooked_ai_b200.mdp import OvercookedGridworld

        mdp_gen_params = self.mdp_params_generator.generate(outside_information)
        if mdp_gen_params.get("layout_name") is not None:
            return self.padded_mdp(OvercookedGridworld.from_layout_name(**mdp_gen_params))
        required_keys = ["inner_shape", "prop_empty", "prop_feats", "display"]
        if not mdp_gen_params.get("generate_all_orders"):
            required_keys.append("start_all_orders")
        missing_keys = [k for k in required_keys if k not in mdp_gen_params]
        assert len(missing_keys) == 0, "These keys were missing from the mdp_params: {}".format(missing_keys)
        inner_shape = mdp_gen_params["inner_shape"]
        assert inner_shape[0] <= self.outer_shape[0] and inner_shape[1] <= self.outer_shape[1], \
            "inner_shape cannot fit into the outershap"
        if "feature_types" not in mdp_gen_params:
            mdp_gen_params["feature_types"] = DEFAULT_FEATURE_TYPES  # written into the caller's dict, as the reference does
        return self.make_new_layout(mdp_gen_params)

    @staticmethod
    def create_base_params(mdp_gen_params):
        """:197-215 — the recipe part of the MDP parameters."""
        assert mdp_gen_params.get("start_all_orders") or mdp_gen_params.get("generate_all_orders")
        mdp_gen_params = LayoutGenerator.add_generated_mdp_params_orders(mdp_gen_params)
        recipe_params = {"start_all_orders": mdp_gen_params["start_all_orders"]}
        if mdp_gen_params.get("start_bonus_orders"):
            recipe_params["start_bonus_orders"] = mdp_gen_params["start_bonus_orders"]
        for k in ("recipe_values", "recipe_times"):
            if k in mdp_gen_params:
                recipe_params[k] = mdp_gen_params[k]
        return recipe_params

    @staticmethod
    def add_generated_mdp_params_orders(mdp_params):
        """:217-254"""
        mdp_params = copy.deepcopy(mdp_params)
        all_recipes = None
        if mdp_params.get("generate_all_orders"):
            kwargs = copy.deepcopy(mdp_params["generate_all_orders"])
            if kwargs.get("recipes"):
                kwargs["recipes"] = [Recipe.from_dict(r) for r in kwargs["recipes"]]
            all_recipes = generate_random_recipes(**kwargs)
            mdp_params["start_all_orders"] = [r.to_dict() for r in all_recipes]
        if mdp_params.get("generate_bonus_orders"):
            kwargs = copy.deepcopy(mdp_params["generate_bonus_orders"])
            if not kwargs.get("recipes"):
                kwargs["recipes"] = all_recipes
            mdp_params["start_bonus_orders"] = [r.to_dict() for r in generate_random_recipes(**kwargs)]
        return mdp_params

    def padded_mdp(self, mdp, display=False):
        """:256-266 — an existing MDP's terrain embedded in the outer shape with NEW random start cells.  Like the
        reference, the result is built from the grid alone: orders and recipe values fall back to the defaults."""
        from overcooked_ai_b200.mdp import OvercookedGridworld

        terrain = np.array([list(row) for row in mdp.terrain_mtx]).T  # [x][y]
        padded = embed(terrain, self.outer_shape)
        starts = random_start_cells(padded)
        return OvercookedGridworld.from_grid(self.padded_grid_to_layout_grid(padded, starts, display=display))

    def make_new_layout(self, mdp_gen_params):
        """:268-276"""
        return self.make_disjoint_sets_layout(
            inner_shape=mdp_gen_params["inner_shape"], prop_empty=mdp_gen_params["prop_empty"],
            prop_features=mdp_gen_params["prop_feats"], base_param=LayoutGenerator.create_base_params(mdp_gen_params),
            feature_types=mdp_gen_params["feature_types"], display=mdp_gen_params["display"])

    def make_disjoint_sets_layout(self, inner_shape, prop_empty, prop_features, base_param,
                                  feature_types=DEFAULT_FEATURE_TYPES, display=True):
        """:278-304"""
        from overcooked_ai_b200.mdp import OvercookedGridworld

        grid = dig_connected_floor(inner_shape, prop_empty)
        place_features(grid, prop_features, feature_types)
        padded = embed(grid, self.outer_shape)
        starts = random_start_cells(padded)
        return OvercookedGridworld.from_grid(self.padded_grid_to_layout_grid(padded, starts, display=display), base_param)

    def padded_grid_to_layout_grid(self, padded_grid, start_positions, display=False):
        if display:
            print("Generated layout")
            print("\n".join(" ".join(str(padded_grid[x, y]) for x in range(padded_grid.shape[0])) for y in range(padded_grid.shape[1])))
        return to_layout_rows(padded_grid, start_positions)


def generate_layout_pool(n, mdp_gen_params=None, outer_shape=(5, 4), skip_unsupported=False):
    """``n`` generated layouts (CompiledLayout) for ``BatchedOvercookedEnv(pool, ..., random_layout=True)``.
    Draws from numpy's global generator like the reference: seed it first for a reproducible pool.
    A draw the packed record cannot hold (more than OVC_MAX_POTS pots — possible with a high ``prop_feats``)
    raises ValueError, or is dropped and redrawn with ``skip_unsupported`` (which biases the pool away from them)."""
    fn = LayoutGenerator.mdp_gen_fn_from_dict(copy.deepcopy(mdp_gen_params or DEFAULT_MDP_GEN_PARAMS), outer_shape=outer_shape)
    pool = []
    while len(pool) < n:
        try:
            pool.append(fn({}).compiled)
        except ValueError:
            if not skip_unsupported:
                raise
    return pool
