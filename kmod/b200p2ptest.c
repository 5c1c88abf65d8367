// SPDX-License-Identifier: GPL-2.0 OR MIT
/*
 * b200p2ptest - kernel half of the B200 GPU P2P test harness.
 *
 * Exposes the NVIDIA P2P page-table interface to user space through /dev/b200p2ptest so that the
 * GPU-driver half of GPUDirect RDMA can be validated with NO NIC and NO OFED: pin a CUDA
 * allocation, read back the bus addresses, map them into the CPU through the GPU's BAR and poke
 * at HBM from a normal process.  Counterpart of the reference's amdp2ptest module
 * (/root/reference/tests/amdp2ptest.c), same session model:
 *
 *   reference                                       here
 *   open :93-112, per-fd list + mutex               b200p2ptest_open
 *   release :115-139, unpin everything on close     b200p2ptest_release
 *   ioctl_is_gpu_address :141-165                   ioctl_is_gpu_address (probe pin: nv-p2p has no classifier)
 *   ioctl_get_page_size :168-205                    ioctl_get_page_size
 *   ioctl_get_pages :207-260                        ioctl_get_pages (no leaked node on any error path)
 *   ioctl_put_pages :263-304, all matches           ioctl_put_pages (reports how many matched)
 *   handlers[] + dispatcher :307-333                b200p2ptest_handlers + b200p2ptest_unlocked_ioctl
 *   mmap :336-395, FIRST sg entry only              b200p2ptest_mmap maps EVERY pinned page of the window
 *   free_callback :77-89                            b200p2ptest_free_callback
 *   misc device, mode 0777 :410-428                 misc device, mode 0660
 *   init/exit :430-474                              b200p2ptest_init / b200p2ptest_exit
 * Extra: GET_BUS_ADDRS returns the bus addresses so a test can compare them with what an HCA
 * would be given.  Unlike the reference the list is never walked without its lock.
 */
#include <linux/module.h>
#include <linux/kernel.h>
#include <linux/slab.h>
#include <linux/mutex.h>
#include <linux/list.h>
#include <linux/fs.h>
#include <linux/miscdevice.h>
#include <linux/uaccess.h>
#include <linux/mm.h>
#include <linux/errno.h>

#include <linux/kref.h>
#include "nv-p2p.h"
#include "b200p2ptest.h"

MODULE_LICENSE("Dual MIT/GPL");
MODULE_DESCRIPTION("NVIDIA B200 GPU P2P basic API test kernel-mode driver");
MODULE_VERSION("1.0");
MODULE_SOFTDEP("pre: nvidia");

#define MSG_INFO(fmt, args...) pr_info(B200P2PTEST_DEVICE_NAME ": " fmt, ##args)
#define MSG_ERR(fmt, args...) pr_err(B200P2PTEST_DEVICE_NAME ": " fmt, ##args)
#define MSG_WARN(fmt, args...) pr_warn(B200P2PTEST_DEVICE_NAME ": " fmt, ##args)

struct b200p2ptest_list;

/* One live pin.  `revoked` is set by the driver's free callback; a revoked node stays on the list
 * (so PUT_PAGES / close still find and free it) but is never unpinned or mapped again. */
struct b200p2ptest_node {
	struct list_head list_node;
	struct b200p2ptest_list *owner;
	u64 handle;
	u64 va;
	u64 size;
	int revoked;
	int zombie; /* PUT_PAGES / close lost a race with a revoke: the free callback, still to come, frees the node */
	struct nvidia_p2p_page_table *page_table;
};

/* Per-open-file state.  Reference counted: the file holds one reference, every free callback that is running
 * holds one, so close() can never free the list (and its mutex) under a callback that is still inside it. */
struct b200p2ptest_list {
	struct list_head head;
	struct mutex lock;
	struct kref ref;
	u64 next_handle;
};

static void b200p2ptest_list_free(struct kref *ref)
{
	struct b200p2ptest_list *list = container_of(ref, struct b200p2ptest_list, ref);

	mutex_destroy(&list->lock);
	kfree(list);
}

static unsigned long node_page_size(const struct b200p2ptest_node *n)
{
	if (n->page_table && n->page_table->page_size == NVIDIA_P2P_PAGE_SIZE_4KB)
		return 4096;
	if (n->page_table && n->page_table->page_size == NVIDIA_P2P_PAGE_SIZE_128KB)
		return 128 * 1024;
	return B200P2P_GPU_PAGE_SIZE;
}

/* The GPU memory behind a pin is going away.  Logged loudly (as the reference does, at ERR level:
 * tests/amdp2ptest.c:81-82) because a tester wants to see it. */
static void b200p2ptest_free_callback(void *data)
{
	struct b200p2ptest_node *node = data;
	struct b200p2ptest_list *list;
	struct nvidia_p2p_page_table *pt;
	int zombie;

	if (!node)
		return;
	/* a callback only fires for a pin that exists, i.e. a node that has not been put yet: node and list are
	 * alive here; the reference keeps the list so past the point where the node may be freed */
	list = node->owner;
	kref_get(&list->ref);
	MSG_ERR("free callback: va 0x%llx size 0x%llx was revoked by the GPU driver\n",
		(unsigned long long)node->va, (unsigned long long)node->size);
	mutex_lock(&list->lock);
	node->revoked = 1;
	pt = node->page_table;
	node->page_table = NULL;
	zombie = node->zombie;
	mutex_unlock(&list->lock);
	/* a live node may be freed by a racing PUT_PAGES / close from here on; a zombie is ours alone */
	if (pt)
		nvidia_p2p_free_page_table(pt); /* never put_pages after a revoke */
	if (zombie) {
		kfree(node);
		kref_put(&list->ref, b200p2ptest_list_free); /* the reference the zombie held */
	}
	kref_put(&list->ref, b200p2ptest_list_free);
}

static int b200p2ptest_open(struct inode *inode, struct file *filp)
{
	struct b200p2ptest_list *list = kzalloc(sizeof(*list), GFP_KERNEL);

	if (!list)
		return -ENOMEM;
	INIT_LIST_HEAD(&list->head);
	mutex_init(&list->lock);
	kref_init(&list->ref);
	list->next_handle = 1;
	filp->private_data = list;
	MSG_INFO("open: session %p\n", list);
	return 0;
}

/* Unlink every node under the lock, then release outside it (put_pages may sleep / call back). */
static int drop_nodes(struct b200p2ptest_list *list, int match, u64 va, u64 size)
{
	struct b200p2ptest_node *node, *tmp;
	struct list_head doomed;
	int n = 0;

	INIT_LIST_HEAD(&doomed);
	mutex_lock(&list->lock);
	list_for_each_entry_safe(node, tmp, &list->head, list_node) {
		if (match && (node->va != va || node->size != size))
			continue;
		list_del(&node->list_node);
		list_add(&node->list_node, &doomed);
	}
	mutex_unlock(&list->lock);
	list_for_each_entry_safe(node, tmp, &doomed, list_node) {
		struct nvidia_p2p_page_table *pt;
		int revoked;

		/* the free callback may still race for this node's page table: same lock decides */
		mutex_lock(&list->lock);
		pt = node->page_table;
		node->page_table = NULL;
		revoked = node->revoked;
		mutex_unlock(&list->lock);
		if (pt && !revoked) {
			int ret = nvidia_p2p_put_pages(0, 0, node->va, pt);

			if (ret) {
				/* Refused: a revoke of this pin is in flight.  If its callback has already been here it
				 * found the page table taken, so releasing it falls to us; if it is still to come, the
				 * node (and a reference on the list it points to) must outlive this call: hand both,
				 * and the page table, to the callback. */
				int handed_over = 0;

				MSG_ERR("put_pages(0x%llx) failed: %d (revoked meanwhile)\n", (unsigned long long)node->va, ret);
				mutex_lock(&list->lock);
				if (!node->revoked) {
					list_del(&node->list_node); /* off our private list BEFORE the callback may free it */
					node->page_table = pt;
					node->zombie = 1;
					kref_get(&list->ref);
					handed_over = 1;
				}
				mutex_unlock(&list->lock);
				if (handed_over) {
					++n; /* node is the callback's from here on: not touched again */
					continue;
				}
				nvidia_p2p_free_page_table(pt);
			}
		}
		list_del(&node->list_node);
		kfree(node);
		++n;
	}
	return n;
}

static int b200p2ptest_release(struct inode *inode, struct file *filp)
{
	struct b200p2ptest_list *list = filp->private_data;
	int n;

	if (!list)
		return 0;
	n = drop_nodes(list, 0, 0, 0);
	MSG_INFO("close: session %p, released %d pin(s) the application left behind\n", list, n);
	filp->private_data = NULL;
	kref_put(&list->ref, b200p2ptest_list_free); /* a free callback still inside the list keeps it alive */
	return 0;
}

/* nv-p2p has no is_gpu_address(): a page is GPU memory iff the driver agrees to pin it. */
static void probe_free_callback(void *data)
{
	WRITE_ONCE(*(int *)data, 1);
}

static int probe_gpu_page(u64 addr)
{
	struct nvidia_p2p_page_table *pt = NULL;
	u64 page = addr & ~(B200P2P_GPU_PAGE_SIZE - 1);
	int revoked = 0;
	int ret = nvidia_p2p_get_pages(0, 0, page, B200P2P_GPU_PAGE_SIZE, &pt, probe_free_callback, &revoked);

	if (ret || !pt)
		return 0;
	if (READ_ONCE(revoked)) {
		nvidia_p2p_free_page_table(pt); /* freed while we looked: no longer GPU memory */
		return 0;
	}
	nvidia_p2p_put_pages(0, 0, page, pt);
	return 1;
}

static long ioctl_is_gpu_address(struct file *filp, unsigned long arg)
{
	struct b200p2p_is_gpu_address p;

	if (copy_from_user(&p, (void __user *)arg, sizeof(p)))
		return -EFAULT;
	p.ret_value = probe_gpu_page(p.addr);
	p.reserved = 0;
	MSG_INFO("IS_GPU_ADDRESS: addr 0x%llx -> %u\n", (unsigned long long)p.addr, p.ret_value);
	if (copy_to_user((void __user *)arg, &p, sizeof(p)))
		return -EFAULT;
	return 0;
}

static long ioctl_get_page_size(struct file *filp, unsigned long arg)
{
	struct b200p2p_get_page_size p;

	if (copy_from_user(&p, (void __user *)arg, sizeof(p)))
		return -EFAULT;
	MSG_INFO("GET_PAGE_SIZE: addr 0x%llx length 0x%llx\n", (unsigned long long)p.addr, (unsigned long long)p.length);
	if (!p.length || !probe_gpu_page(p.addr) || !probe_gpu_page(p.addr + p.length - 1))
		return -EFAULT; /* same errno the reference returns when the GPU driver refuses */
	p.page_size = B200P2P_GPU_PAGE_SIZE;
	if (copy_to_user((void __user *)arg, &p, sizeof(p)))
		return -EFAULT;
	return 0;
}

static long ioctl_get_pages(struct file *filp, unsigned long arg)
{
	struct b200p2ptest_list *list = filp->private_data;
	struct b200p2p_get_pages p;
	struct b200p2ptest_node *node;
	struct nvidia_p2p_page_table *pt = NULL, *dead = NULL;
	int ret;

	if (copy_from_user(&p, (void __user *)arg, sizeof(p)))
		return -EFAULT;
	MSG_INFO("GET_PAGES: addr 0x%llx length 0x%llx\n", (unsigned long long)p.addr, (unsigned long long)p.length);
	if (!p.length || (p.addr & (B200P2P_GPU_PAGE_SIZE - 1)) || (p.length & (B200P2P_GPU_PAGE_SIZE - 1)))
		return -EINVAL;
	node = kzalloc(sizeof(*node), GFP_KERNEL);
	if (!node)
		return -ENOMEM;
	node->owner = list;
	node->va = p.addr;
	node->size = p.length;
	/* on the list BEFORE the pin exists, so a revoke that fires immediately finds a live owner */
	mutex_lock(&list->lock);
	node->handle = list->next_handle++;
	list_add(&node->list_node, &list->head);
	mutex_unlock(&list->lock);

	ret = nvidia_p2p_get_pages(0, 0, p.addr, p.length, &pt, b200p2ptest_free_callback, node);
	if (ret || !pt) {
		mutex_lock(&list->lock);
		list_del(&node->list_node);
		mutex_unlock(&list->lock);
		kfree(node);
		return ret ? -EFAULT : -EINVAL;
	}
	/* publish the pin under the lock the free callback takes: it may already have fired (the memory was freed
	 * the instant it was pinned), in which case it found no page table to release and that falls to us */
	mutex_lock(&list->lock);
	if (node->revoked) {
		dead = pt;
	} else {
		node->page_table = pt;
	}
	p.handle = node->handle;
	p.entries = node->page_table ? node->page_table->entries : 0;
	p.page_size = (u32)node_page_size(node);
	mutex_unlock(&list->lock);
	if (dead)
		nvidia_p2p_free_page_table(dead);
	if (copy_to_user((void __user *)arg, &p, sizeof(p))) {
		drop_nodes(list, 1, p.addr, p.length); /* nothing leaks: the reference leaks the node here */
		return -EFAULT;
	}
	return 0;
}

static long ioctl_put_pages(struct file *filp, unsigned long arg)
{
	struct b200p2ptest_list *list = filp->private_data;
	struct b200p2p_put_pages p;

	if (copy_from_user(&p, (void __user *)arg, sizeof(p)))
		return -EFAULT;
	/* every pin of exactly this range goes: "to allow test situation when get_pages would be called
	 * on the same memory several times" (tests/amdp2ptest.c:296-299) */
	p.released = (u32)drop_nodes(list, 1, p.addr, p.length);
	p.reserved = 0;
	MSG_INFO("PUT_PAGES: addr 0x%llx length 0x%llx -> %u pin(s) released\n", (unsigned long long)p.addr,
		 (unsigned long long)p.length, p.released);
	if (copy_to_user((void __user *)arg, &p, sizeof(p)))
		return -EFAULT;
	return 0;
}

static long ioctl_get_bus_addrs(struct file *filp, unsigned long arg)
{
	struct b200p2ptest_list *list = filp->private_data;
	struct b200p2p_get_bus_addrs *p;
	struct b200p2ptest_node *node;
	long ret = -ENOENT;
	u32 i, n = 0;

	p = kzalloc(sizeof(*p), GFP_KERNEL);
	if (!p)
		return -ENOMEM;
	if (copy_from_user(p, (void __user *)arg, sizeof(*p))) {
		kfree(p);
		return -EFAULT;
	}
	if (p->count > B200P2P_MAX_BUS_ADDRS)
		p->count = B200P2P_MAX_BUS_ADDRS;
	mutex_lock(&list->lock);
	list_for_each_entry(node, &list->head, list_node) {
		if (node->handle != p->handle)
			continue;
		if (node->revoked || !node->page_table) {
			ret = -ESTALE;
			break;
		}
		for (i = p->first; i < node->page_table->entries && n < p->count; ++i)
			p->addrs[n++] = node->page_table->pages[i]->physical_address;
		ret = 0;
		break;
	}
	mutex_unlock(&list->lock);
	p->count = n;
	if (!ret && copy_to_user((void __user *)arg, p, sizeof(*p)))
		ret = -EFAULT;
	kfree(p);
	return ret;
}

static const struct {
	unsigned int cmd;
	long (*handler)(struct file *filp, unsigned long arg);
} b200p2ptest_handlers[] = {
	{ B200P2PTEST_IOCTL_IS_GPU_ADDRESS, ioctl_is_gpu_address },
	{ B200P2PTEST_IOCTL_GET_PAGE_SIZE, ioctl_get_page_size },
	{ B200P2PTEST_IOCTL_GET_PAGES, ioctl_get_pages },
	{ B200P2PTEST_IOCTL_PUT_PAGES, ioctl_put_pages },
	{ B200P2PTEST_IOCTL_GET_BUS_ADDRS, ioctl_get_bus_addrs },
};

static long b200p2ptest_unlocked_ioctl(struct file *filp, unsigned int cmd, unsigned long arg)
{
	unsigned int i;

	for (i = 0; i < ARRAY_SIZE(b200p2ptest_handlers); i++)
		if (cmd == b200p2ptest_handlers[i].cmd)
			return b200p2ptest_handlers[i].handler(filp, arg);
	return -EINVAL;
}

/*
 * CPU window onto pinned GPU memory: mmap(fd, len, ..., offset = GPU VA).  The window must lie inside
 * one live pin; every 64 KiB GPU page of it is mapped at its own bus address (the reference maps the
 * whole window from the FIRST scatterlist entry and returns inside the loop: tests/amdp2ptest.c:372-390).
 */
static int b200p2ptest_mmap(struct file *filp, struct vm_area_struct *vma)
{
	struct b200p2ptest_list *list = filp->private_data;
	u64 gpu_va = (u64)vma->vm_pgoff << PAGE_SHIFT;
	u64 size = vma->vm_end - vma->vm_start;
	struct b200p2ptest_node *node;
	int ret = -EINVAL;

	MSG_INFO("mmap: GPU va 0x%llx size 0x%llx\n", (unsigned long long)gpu_va, (unsigned long long)size);
	if (!size || (gpu_va & (PAGE_SIZE - 1)))
		return -EINVAL;
	mutex_lock(&list->lock);
	list_for_each_entry(node, &list->head, list_node) {
		unsigned long psz, user = vma->vm_start;
		u64 off, left = size;

		if (node->revoked || !node->page_table)
			continue;
		if (gpu_va < node->va || gpu_va + size > node->va + node->size)
			continue;
		psz = node_page_size(node);
		off = gpu_va - node->va;
		vm_flags_set(vma, VM_IO | VM_PFNMAP | VM_DONTEXPAND | VM_DONTDUMP);
		vma->vm_page_prot = pgprot_writecombine(vma->vm_page_prot);
		ret = 0;
		while (left) {
			u64 idx = off / psz, in_page = off % psz;
			u64 chunk = min_t(u64, left, psz - in_page);
			u64 bus = node->page_table->pages[idx]->physical_address + in_page;

			ret = io_remap_pfn_range(vma, user, bus >> PAGE_SHIFT, chunk, vma->vm_page_prot);
			if (ret) {
				MSG_ERR("mmap: remap of GPU page %llu failed: %d\n", (unsigned long long)idx, ret);
				break;
			}
			user += chunk;
			off += chunk;
			left -= chunk;
		}
		break;
	}
	mutex_unlock(&list->lock);
	return ret;
}

static const struct file_operations b200p2ptest_fops = {
	.owner = THIS_MODULE,
	.open = b200p2ptest_open,
	.release = b200p2ptest_release,
	.unlocked_ioctl = b200p2ptest_unlocked_ioctl,
	.mmap = b200p2ptest_mmap,
};

static struct miscdevice b200p2ptest_dev = {
	.minor = MISC_DYNAMIC_MINOR,
	.name = B200P2PTEST_DEVICE_NAME,
	.fops = &b200p2ptest_fops,
	.mode = S_IRUSR | S_IWUSR | S_IRGRP | S_IWGRP, /* not world-writable: it hands out bus addresses */
};

static int __init b200p2ptest_init(void)
{
	int ret = misc_register(&b200p2ptest_dev);

	if (ret) {
		MSG_ERR("cannot register the misc device: %d\n", ret);
		return ret;
	}
	/* load-time smoke check, as the reference prints its four KFD entry points (tests/amdp2ptest.c:441-445):
	 * the addresses the NVIDIA P2P symbols resolved to -- a NULL here means nvidia.ko is not what we linked against */
	MSG_INFO("nvidia_p2p_get_pages %p put_pages %p free_page_table %p dma_map_pages %p\n", (void *)nvidia_p2p_get_pages,
		 (void *)nvidia_p2p_put_pages, (void *)nvidia_p2p_free_page_table, (void *)nvidia_p2p_dma_map_pages);
	MSG_INFO("ready: %s, ABI %d, GPU page %llu KiB\n", B200P2PTEST_DEVICE_PATH, B200P2PTEST_ABI_VERSION,
		 (unsigned long long)(B200P2P_GPU_PAGE_SIZE >> 10));
	return 0;
}

static void __exit b200p2ptest_exit(void)
{
	misc_deregister(&b200p2ptest_dev);
	MSG_INFO("removed\n");
}

module_init(b200p2ptest_init);
module_exit(b200p2ptest_exit);
