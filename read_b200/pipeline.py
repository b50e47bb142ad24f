"""Drop-in for ``READ.pipelines.ogl.TexturePipeline`` (READ/pipelines/ogl.py:58-154) over the B200 kernels.

Use ``--pipeline read_b200.pipeline.TexturePipeline`` with the reference's train.py (the plugin is located
by dotted path, READ/utils/train.py:148-154), or shadow ``READ.pipelines.ogl``.  Same CLI flags,
same attributes after ``create`` (model, ds_train, ds_val, optimizer, criterion, net, textures), same
``state_objects`` / ``dataset_load`` / ``dataset_unload`` / ``extra_optimizer`` / ``get_net`` contract, same
checkpoint format ({'state_dict','args'}, READ/utils/train.py:42-65).

Dataset construction (``get_datasets``) and the loss stay the reference's own code: they are outside the render
hot path and are imported lazily from ``READ`` only in training mode.
"""
from pathlib import Path

import torch
from torch import optim

from .texture import PointTexture
from .unet import UNet
from .compose import NetAndTexture

TextureOptimizerClass = optim.RMSprop        # ogl.py:16


class Pipeline:
    """READ/pipelines/pipeline.py:10-31."""

    def export_args(self, parser):
        raise NotImplementedError()

    def create(self, args):
        raise NotImplementedError()

    def dataset_load(self, *args, **kwargs):
        pass

    def dataset_unload(self, *args, **kwargs):
        pass

    def get_net(self):
        raise NotImplementedError()

    def extra_optimizer(self, *args):
        return None


def load_model_checkpoint(path, model):
    """READ/utils/train.py:60-65."""
    ckpt = torch.load(path, map_location='cpu')
    model.load_state_dict(ckpt['state_dict'])
    return model


def save_model(save_path, model, args=None):
    """READ/utils/train.py:42-57 ({'state_dict', 'args'})."""
    m = model.module if hasattr(model, 'module') else model
    d = {'state_dict': m.state_dict()}
    if args is not None:
        d['args'] = dict(vars(args)) if hasattr(args, '__dict__') else dict(args)
    torch.save(d, save_path)


def get_net(input_channels, args):
    return UNet(num_input_channels=8, num_output_channels=3, feature_scale=4, num_res=4)   # ogl.py:19-27


def get_texture(num_channels, size, args):
    if not hasattr(args, 'reg_weight'):
        args.reg_weight = 0.
    if getattr(args, 'use_mesh', False):
        raise NotImplementedError("read_b200: mesh textures (MeshTexture) are outside the point-cloud hot path")
    texture = PointTexture(num_channels, size, activation=args.texture_activation, reg_weight=args.reg_weight)
    if getattr(args, 'texture_ckpt', None):
        texture = load_model_checkpoint(args.texture_ckpt, texture)
    return texture


class TexturePipeline(Pipeline):
    def export_args(self, parser):
        add = getattr(parser, 'add', parser.add_argument)
        parser.add_argument('--descriptor_size', type=int, default=8)
        parser.add_argument('--texture_size', type=int)
        parser.add_argument('--texture_ckpt', type=Path)
        add('--texture_lr', type=float, default=1e-1)
        add('--texture_activation', type=str, default='none')
        add('--n_points', type=int, default=0, help='this is for inference')

    def create(self, args):
        if not hasattr(args, 'input_channels'):
            args.input_channels = None
        if not args.input_channels:
            args.input_channels = [args.descriptor_size] * getattr(args, 'num_mipmap', 5)
        net = get_net(args.input_channels, args)
        textures = {}
        if getattr(args, 'inference', False):
            textures = {0: get_texture(args.descriptor_size, args.n_points, args)}
        else:
            from READ.datasets.dynamic import get_datasets          # reference data path, out of scope here
            self.ds_train, self.ds_val = get_datasets(args)
            for ds in self.ds_train:
                assert ds.scene_data['pointcloud'] is not None, 'set pointcloud'
                size = ds.scene_data['pointcloud']['xyz'].shape[0]
                textures[ds.id] = get_texture(args.descriptor_size, size, args)
            self.optimizer = optim.Adam(net.parameters(), lr=args.lr)
            if len(textures) == 1:
                self._extra_optimizer = TextureOptimizerClass(textures[0].parameters(), lr=args.texture_lr)
            else:
                self._extra_optimizer = None
            self.criterion = args.criterion_module(**args.criterion_args).cuda()
        ss = args.supersampling if hasattr(args, 'supersampling') else 1
        self.net = net
        self.textures = textures
        self.model = NetAndTexture(net, textures, ss)
        self.args = args

    def state_objects(self):
        objs = {'net': self.net}
        objs.update({ds.name: self.textures[ds.id] for ds in self.ds_train})
        return objs

    def dataset_load(self, dataset):
        self.model.load_textures([ds.id for ds in dataset])
        for ds in dataset:
            ds.load()

    def extra_optimizer(self, dataset):
        lr_drop = self.optimizer.param_groups[0]['lr'] / self.args.lr
        if self._extra_optimizer is not None:      # single dataset: keep optimizer state
            self._extra_optimizer.param_groups[0]['lr'] = self.args.texture_lr * lr_drop
            return self._extra_optimizer
        groups = [{'params': self.textures[ds.id].parameters()} for ds in dataset]
        return TextureOptimizerClass(groups, lr=self.args.texture_lr * lr_drop)

    def dataset_unload(self, dataset):
        self.model.unload_textures()
        for ds in dataset:
            ds.unload()
            self.textures[ds.id].null_grad()

    def get_net(self):
        return self.net
