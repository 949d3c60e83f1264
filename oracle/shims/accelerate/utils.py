def broadcast_object_list(objects, from_process=0):
    """World size 1: nothing to broadcast."""
    return objects
