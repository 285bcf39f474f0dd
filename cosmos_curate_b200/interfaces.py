"""The reference's plugin surface (drop-in boundary of this path).

If the reference package is importable (a real cosmos-curate environment: ray + the xenna Rust extension),
its own classes are re-exported and the B200 stages subclass THEM, so they can be listed in
`_assemble_stages` of splitting_pipeline.py unchanged.  Otherwise (this repo's tests, bench) identical
stand-ins are defined here, restating

    PipelineTask / CuratorStageResource / CuratorStage / CuratorStageSpec
        cosmos_curate/core/interfaces/stage_interface.py:27-214
    ModelInterface
        cosmos_curate/core/interfaces/model_interface.py:20-54
    run_pipeline(tasks, stages, runner=...)           (test-mode subset)
        cosmos_curate/core/interfaces/pipeline_interface.py:281-329
    SequentialRunner
        tests/utils/sequential_runner.py:27-69

with the same names, argument meaning and call order (stage_setup() once, then process_data(list[task]) ->
list[task] | None, then destroy()).
"""

from __future__ import annotations

import abc
from collections.abc import Sequence

import attrs

try:  # pragma: no cover - only inside a full cosmos-curate environment
    from cosmos_curate.core.interfaces.model_interface import ModelInterface
    from cosmos_curate.core.interfaces.stage_interface import CuratorStage, CuratorStageResource, CuratorStageSpec, PipelineTask

    USING_REFERENCE_INTERFACES = True
except Exception:  # noqa: BLE001 - ray / cosmos_xenna missing: use the stand-ins
    USING_REFERENCE_INTERFACES = False

    @attrs.define
    class PipelineTask:
        """stage_interface.py:27-56."""

        @property
        def weight(self) -> float:
            return 1.0

        @property
        def fraction(self) -> float:
            return 1.0

        def get_major_size(self) -> int:
            return 0

    @attrs.define
    class CuratorStageResource:
        """stage_interface.py:59-65."""

        cpus: float = 1.0
        gpus: float | int = 0

    class ModelInterface(abc.ABC):
        """model_interface.py:20-54."""

        @property
        @abc.abstractmethod
        def conda_env_name(self) -> str: ...

        @property
        @abc.abstractmethod
        def model_id_names(self) -> list[str]: ...

        @abc.abstractmethod
        def setup(self) -> None: ...

    class CuratorStage:
        """stage_interface.py:69-188 (without the xenna / ray plumbing)."""

        @property
        def resources(self) -> CuratorStageResource:
            return CuratorStageResource(cpus=1.0, gpus=0.0)

        @property
        def model(self) -> ModelInterface | None:
            return None

        @property
        def conda_env_name(self) -> str | None:
            return self.model.conda_env_name if self.model is not None else None

        def stage_setup_on_node(self) -> None:
            return

        def stage_setup(self) -> None:
            if self.model is not None:
                self.model.setup()

        def process_data(self, task: list[PipelineTask]) -> list[PipelineTask] | None:
            return task

        def destroy(self) -> None:
            return

        @property
        def stage_batch_size(self) -> int:
            return 1

        # same plumbing names as the reference; no-ops without xenna
        def setup_on_node(self, _node_info=None, _worker_metadata=None) -> None:
            self.stage_setup_on_node()

        def setup(self, _=None) -> None:
            self.stage_setup()

    @attrs.define
    class CuratorStageSpec:
        """stage_interface.py:191-214 / xenna StageSpec fields used by this path."""

        stage: CuratorStage
        num_workers_per_node: int | None = None
        num_run_attempts_python: int = 1
        over_provision_factor: float | None = None
        worker_max_lifetime_m: int = 0

        def name(self) -> str:
            return self.stage.__class__.__name__

        def display_str(self) -> str:
            res = self.name()
            res += f" num_workers_per_node={self.num_workers_per_node}"
            res += f" cpus={self.stage.resources.cpus}"
            res += f" gpus={self.stage.resources.gpus}"
            return res


class SequentialRunner:
    """tests/utils/sequential_runner.py:27-69: setup every stage, then per stage process_data -> destroy."""

    def run(self, input_tasks, stage_specs: Sequence, _model_weights_prefix: str = "", _execution_mode: str = "AUTO"):
        stages = [spec.stage for spec in stage_specs]
        for stage in stages:
            stage.stage_setup()
        tasks = input_tasks
        for stage in stages:
            result = stage.process_data(tasks)
            if result is None:
                return None
            tasks = result
            stage.destroy()
        return tasks


def run_pipeline(input_tasks, stages: Sequence, runner=None, model_weights_prefix: str = ""):
    """Test-mode subset of pipeline_interface.run_pipeline: wrap bare stages in specs, hand them to the runner."""
    specs = [s if isinstance(s, CuratorStageSpec) else CuratorStageSpec(s) for s in stages]
    runner = runner or SequentialRunner()
    return runner.run(input_tasks, specs, model_weights_prefix)
