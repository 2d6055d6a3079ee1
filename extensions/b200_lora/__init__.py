# ai-toolkit extension: copy or symlink this directory into <ai-toolkit>/extensions/ (and put this repository on
# PYTHONPATH), then select it in a job config with   process: - type: sd_trainer_b200   (everything else unchanged).
# Registry contract: toolkit/extension.py:9-57, extensions/example/__init__.py.
from toolkit.extension import Extension


class B200LoRATrainerExtension(Extension):
    uid = "sd_trainer_b200"
    name = "SD Trainer (B200 fused LoRA path)"

    @classmethod
    def get_process(cls):
        # imports stay in here so that scanning the extensions does not load CUDA code (same rule as the example extension)
        from extensions_built_in.sd_trainer.SDTrainer import SDTrainer
        from toolkit.scheduler import get_lr_scheduler

        from ai_toolkit_b200.plugin import make_trainer_class

        return make_trainer_class(SDTrainer, get_lr_scheduler)


AI_TOOLKIT_EXTENSIONS = [B200LoRATrainerExtension]
