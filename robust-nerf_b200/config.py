"""Configuration dataclasses -- same names, fields and defaults as noisy_src/config.py:10-131 so
callers can pass either the reference's objects or these (only attribute access is used)."""
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional, Tuple


@dataclass
class ModelConfig:
    pos_freqs: int = 10
    dir_freqs: int = 4
    hidden_dim: int = 256
    num_hidden_layers: int = 8
    skips: Tuple[int, ...] = (4,)
    use_view_dirs: bool = True


@dataclass
class RenderConfig:
    near: float = 2.0
    far: float = 6.0
    num_samples: int = 64
    num_samples_fine: int = 128
    use_hierarchical: bool = True
    perturb: bool = True
    raw_noise_std: float = 0.0
    white_background: bool = True


@dataclass
class DataConfig:
    scene_name: str = "lego"
    data_root: Optional[Path] = None
    img_scale: float = 0.5
    batch_size: int = 1024
    shuffle: bool = True


@dataclass
class TrainConfig:
    lr: float = 5e-4
    lr_decay: int = 250
    num_iterations: int = 200000
    log_every: int = 100
    save_every: int = 10000
    val_every: int = 5000
    output_dir: Path = field(default_factory=lambda: Path("outputs"))
    experiment_name: str = "baseline"
    device: str = "cuda"
    seed: int = 42


@dataclass
class PoseOptConfig:
    enabled: bool = True
    learn_rotation: bool = True
    learn_translation: bool = True
    pose_lr: float = 1e-4
    pose_opt_delay: int = 1000
    init_mode: str = "noisy"
    rotation_noise_deg: float = 0.0
    translation_noise_pct: float = 0.0
    noise_seed: Optional[int] = None


@dataclass
class NeRFConfig:
    model: ModelConfig = field(default_factory=ModelConfig)
    render: RenderConfig = field(default_factory=RenderConfig)
    data: DataConfig = field(default_factory=DataConfig)
    train: TrainConfig = field(default_factory=TrainConfig)
    pose_opt: Optional[PoseOptConfig] = None

    def __post_init__(self):
        if isinstance(self.train.output_dir, str):
            self.train.output_dir = Path(self.train.output_dir)
        if isinstance(self.data.data_root, str):
            self.data.data_root = Path(self.data.data_root)
