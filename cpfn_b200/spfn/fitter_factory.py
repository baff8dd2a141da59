"""Primitive-type registry with the function names of the reference ``SPFN/fitter_factory.py:5-30``.

Ids follow the order passed to ``register_primitives`` (the reference configs register
sphere=0, plane=1, cylinder=2, cone=3, ``Configs/*.yml:13-17``).  ``primitive_name_to_id_dict`` stays a
module global, as callers of the reference read it directly.
"""
KNOWN_TYPES = ("plane", "sphere", "cylinder", "cone")
primitive_name_to_id_dict = {}


def register_primitives(primitive_name_list):
    """Must run once before anything asks for an id."""
    global primitive_name_to_id_dict
    primitive_name_to_id_dict = {name: i for i, name in enumerate(primitive_name_list)}
    print('Registered ' + ','.join(primitive_name_list))


def primitive_name_to_id(name):
    return primitive_name_to_id_dict[name]


def get_n_registered_primitives():
    return len(primitive_name_to_id_dict)


def create_primitive_from_dict(d):
    """The reference builds numpy value objects (``SPFN/primitives.py``) here; those classes are
    host-side metadata outside the hot path, so the validated dictionary itself is returned."""
    if d['type'] not in KNOWN_TYPES:
        raise NotImplementedError
    return dict(d)
