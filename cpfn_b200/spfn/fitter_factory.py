"""Primitive-name registry with the functions of the reference SPFN/fitter_factory.py:5-30
(ids follow the order passed to ``register_primitives``; the reference configs use
sphere=0, plane=1, cylinder=2, cone=3)."""
primitive_name_to_id_dict = {}


def primitive_name_to_id(name):
    return primitive_name_to_id_dict[name]


def get_n_registered_primitives():
    return len(primitive_name_to_id_dict)


def register_primitives(primitive_name_list):
    global primitive_name_to_id_dict
    primitive_name_to_id_dict = {}
    for idx, name in enumerate(primitive_name_list):
        primitive_name_to_id_dict[name] = idx
    print('Registered ' + ','.join(primitive_name_list))


def create_primitive_from_dict(d):
    """The reference builds numpy value objects (SPFN/primitives.py) here; those classes are
    host-side metadata outside the hot path, so this returns the validated dict itself."""
    if d['type'] not in ('plane', 'sphere', 'cylinder', 'cone'):
        raise NotImplementedError
    return dict(d)
