"""Host helpers of the hot path, same names / arguments / return conventions as the reference's
`utils/utils.py` (SaveHelper :10-49, colormap :52-77, calculate_miou :80-126, mini_batch :129-185,
string_class_iou :188-214, choose_frames :237-254).  NumPy only; nothing here touches TensorFlow."""
import random
from collections import deque

import numpy as np

try:                                    # OpenCV is only needed when mini_batch has to rescale
    import cv2
except ImportError:                     # pragma: no cover
    cv2 = None


class SaveHelper:
    """name -> tensor get/set against the device-resident student (reference: placeholder + tf.assign per global
    variable, utils/utils.py:12-18).  `map_fun` filters names exactly as in the reference: a key is transferred
    iff map_fun(key) is not None; a key the network does not have raises KeyError (utils/utils.py:41)."""

    def __init__(self, student, map_fun=lambda x: x):
        self.student = student
        self.map_fun = map_fun

    def save_vars(self, sess, vars_list, map_fun, save_dir=None):
        save_dict = {}
        for name in vars_list:
            key = map_fun(name)
            if key:
                save_dict[key] = self.student.get_tensor(name)
        if save_dir:
            np.save(save_dir, save_dict)
        return save_dict

    def restore_vars(self, sess, load_dir, map_fun):
        print('Trying to restore checkpoint')
        if isinstance(load_dir, str):
            vars_list = np.load(load_dir, allow_pickle=True).item()
        elif isinstance(load_dir, dict):
            vars_list = load_dir
        else:
            exit(1)
        for var_name in vars_list:
            if map_fun(var_name) is not None:
                self.student.set_tensor(var_name, vars_list[var_name])
        print('Restored successfully')
        return


def colormap(name='cityscapes'):
    if name != 'cityscapes':
        raise Exception('Unknown colormap')
    cm = np.zeros((256, 3), dtype=np.uint8)
    cm[:19] = [[128, 64, 128], [244, 35, 232], [70, 70, 70], [102, 102, 156], [190, 153, 153], [153, 153, 153],
               [250, 170, 30], [220, 220, 0], [107, 142, 35], [152, 251, 152], [70, 130, 180], [220, 20, 60],
               [255, 0, 0], [0, 0, 142], [0, 0, 70], [0, 60, 100], [0, 80, 100], [0, 0, 230], [119, 11, 32]]
    return cm


def calculate_miou(conf_matrix, population=False, detailed=False, nan=False):
    """Per-class IoU from a confusion matrix (rows = labels, cols = predictions): cm[i,i] / (row_i + col_i - cm[i,i]);
    a class that is neither present nor predicted gives NaN (nan=True) or the string 'Not predicted/present'."""
    cm = np.asarray(conf_matrix)
    n = len(cm[0])
    miou, false_pos, false_neg = [], [], []
    for i in range(n):
        denominator = cm[i, :].sum() + cm[:, i].sum() - cm[i][i]
        if denominator == 0:
            miou.append(np.nan if nan else 'Not predicted/present')
            if detailed:
                false_pos.append(0)
                false_neg.append(0)
        else:
            miou.append(cm[i][i] / (max(denominator, 1)))
            if detailed:
                false_neg.append((np.sum(cm[i]) - cm[i][i]) / denominator)
                false_pos.append((np.sum(cm[:, i]) - cm[i][i]) / denominator)
    if population:
        population_class = np.sum(cm, axis=1)
        if detailed:
            return miou, population_class / np.sum(population_class), false_neg, false_pos
        return miou, population_class / np.sum(population_class)
    if detailed:
        return miou, false_neg, false_pos
    return miou


def mini_batch(deque_images, deque_labels, crop_size, scale, mini_batch_size, num_of_iterations, flip=False):
    """Random frame + random crop (+ optional rescale / flip) batches, float64 like the reference:
    returns ([iters, B, H, W, 3], [iters, B, H, W]).  RNG draws are made in the reference's order
    (np.random.choice for the frame; random.randint for the scale and the two offsets; np.random.random for flip)."""
    images = list(deque_images) if isinstance(deque_images, deque) else deque_images
    labels = list(deque_labels) if isinstance(deque_labels, deque) else deque_labels
    out_i = np.empty((num_of_iterations, mini_batch_size, crop_size[0], crop_size[1], images[0].shape[2]))
    out_l = np.empty((num_of_iterations, mini_batch_size, crop_size[0], crop_size[1]))
    cache_i = {s: {} for s in scale}
    cache_l = {s: {} for s in scale}
    total = len(images)
    for i in range(num_of_iterations):
        for j in range(mini_batch_size):
            pic = np.random.choice(total)
            h_img, w_img = images[pic].shape[0], images[pic].shape[1]
            chosen = scale[random.randint(0, len(scale) - 1)]
            actual = chosen * crop_size[1] / w_img
            max_h = int(h_img * actual) - crop_size[0]
            max_w = int(w_img * actual) - crop_size[1]
            assert max_w >= 0
            assert max_h >= 0
            h = random.randint(0, max_h)
            w = random.randint(0, max_w)
            if pic not in cache_i[chosen]:
                if actual == 1 and chosen == 1:
                    cache_i[chosen][pic], cache_l[chosen][pic] = images[pic], labels[pic]
                else:
                    size = (int(w_img * actual), int(h_img * actual))
                    cache_i[chosen][pic] = cv2.resize(images[pic], size, interpolation=cv2.INTER_LINEAR)
                    cache_l[chosen][pic] = cv2.resize(labels[pic], size, fx=0, fy=0, interpolation=cv2.INTER_NEAREST)
            img = cache_i[chosen][pic][h:h + crop_size[0], w:w + crop_size[1], :]
            lab = cache_l[chosen][pic][h:h + crop_size[0], w:w + crop_size[1]]
            if flip and np.random.random() > 0.5:
                img, lab = np.flip(img, axis=1), np.flip(lab, axis=1)
            out_i[i][j] = img
            out_l[i][j] = lab
    return out_i, out_l


_CITYSCAPES_NAMES = ['road', 'sidewalk', 'building', 'wall', 'fence', 'pole', 'traffic light', 'traffic sign', 'vegetation',
                     'terrain', 'sky', 'person', 'rider', 'car', 'truck', 'bus', 'train', 'motorcycle', 'bicycle']


def string_class_iou(class_iou_list, population=None, headers=None, class_weights=None):
    out = ''
    if headers is not None:
        out = '%22s\t' % '' + ''.join(h + '\t\t' for h in headers) + '\n'
    names = _CITYSCAPES_NAMES
    if class_weights is not None:
        names = [names[i] for i in np.where(class_weights == 1)[0]]
    if not type(class_iou_list[0]) == list:
        class_iou_list = [class_iou_list]
    for i in range(len(class_iou_list[0])):
        tag = names[i] + ('(%.3g):' % (population[i] * 100.0) if population is not None else ':')
        out += '%-22s\t' % tag
        for col in class_iou_list:
            out += (col[i] + '\t') if type(col[i]) == str else ('%.1f' % (col[i] * 100.0) + '\t\t\t')
        out += '\n'
    return out


def choose_frames(frame_label_list, sample_fraction):
    """Equally spaced samples: round(fraction * len) frames at indices round(linspace(-1, len-1, k+1)[1:])."""
    samples = int(np.round(sample_fraction * len(frame_label_list)))
    indices = np.linspace(-1, len(frame_label_list) - 1, samples + 1, endpoint=True)[1:]
    indices = np.round(indices).astype(int)
    assert indices.size == samples, f"indices had {indices.size} values but samples is {samples}"
    return ([frame_label_list[k][0] for k in indices], [frame_label_list[k][1] for k in indices])
