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
    """The reference's dispatch (SPFN/fitter_factory.py:21-31): the numpy value objects of ``SPFN/primitives.py`` are
    host-side metadata outside the hot path, built by the reference's own ``<type>_fitter.create_primitive_from_dict``
    (forwarded to its files, see _reference.py; needs the reference checkout on sys.path).  The data pipeline calls
    ``get_primitive_name()`` and ``extract_parameter_data_as_dict`` on them (Utils/dataset_utils.py:79-112)."""
    from . import cone_fitter, cylinder_fitter, plane_fitter, sphere_fitter
    makers = {"plane": plane_fitter, "sphere": sphere_fitter, "cylinder": cylinder_fitter, "cone": cone_fitter}
    if d['type'] not in makers:
        raise NotImplementedError
    return makers[d['type']].create_primitive_from_dict(d)
