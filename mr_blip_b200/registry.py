"""Name -> class registry with the call surface of lavis/common/registry.py:9-329 (the subset the
moment-retrieval path touches: models, tasks, runners, lr schedulers, paths)."""


class Registry:
    mapping = {"model_name_mapping": {}, "task_name_mapping": {}, "runner_name_mapping": {},
               "lr_scheduler_name_mapping": {}, "paths": {}, "state": {}}

    @classmethod
    def _register(cls, kind, name):
        def wrap(obj):
            table = cls.mapping[kind]
            if name in table and table[name] is not obj:
                raise KeyError("Name '%s' already registered for %s." % (name, table[name]))
            table[name] = obj
            return obj
        return wrap

    @classmethod
    def register_model(cls, name):
        return cls._register("model_name_mapping", name)

    @classmethod
    def register_task(cls, name):
        return cls._register("task_name_mapping", name)

    @classmethod
    def register_runner(cls, name):
        return cls._register("runner_name_mapping", name)

    @classmethod
    def register_lr_scheduler(cls, name):
        return cls._register("lr_scheduler_name_mapping", name)

    @classmethod
    def register_path(cls, name, path):
        cls.mapping["paths"][name] = path

    @classmethod
    def get_model_class(cls, name):
        return cls.mapping["model_name_mapping"].get(name)

    @classmethod
    def get_task_class(cls, name):
        return cls.mapping["task_name_mapping"].get(name)

    @classmethod
    def get_runner_class(cls, name):
        return cls.mapping["runner_name_mapping"].get(name)

    @classmethod
    def get_lr_scheduler_class(cls, name):
        return cls.mapping["lr_scheduler_name_mapping"].get(name)

    @classmethod
    def get_path(cls, name):
        return cls.mapping["paths"].get(name)

    @classmethod
    def list_models(cls):
        return sorted(cls.mapping["model_name_mapping"].keys())


registry = Registry()
