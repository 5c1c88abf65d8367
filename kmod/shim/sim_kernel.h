/* SPDX-License-Identifier: MIT */
/*
 * Userspace stand-in for the slice of the Linux kernel API that b200p2p.c and b200p2ptest.c use.
 * With -DB200P2P_SIM -Ikmod/shim the two modules compile UNCHANGED into a shared object in which
 * a mock NVIDIA P2P provider and a mock ib_core (sim_runtime.c) play the external layers, so the
 * registration state machine, revocation ordering, leak-proof close and the ioctl ABI are unit
 * tested without loading anything (tests/test_kmod_sim.py).  The reference ships no fake of either
 * external interface; this is the decoupling trick it lacked.
 */
#ifndef B200_SIM_KERNEL_H_
#define B200_SIM_KERNEL_H_

#include <errno.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/ioctl.h>

#define __KERNEL__ 1

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t s64;
typedef unsigned int gfp_t;
typedef uint64_t dma_addr_t;
typedef unsigned long pgprot_t;

#define __init
#define __exit
#define __user
#define __iomem
#define GFP_KERNEL 0u
#define PAGE_SHIFT 12
#define PAGE_SIZE (1UL << PAGE_SHIFT)
#define S_IRUSR 0400
#define S_IWUSR 0200
#define S_IRGRP 0040
#define S_IWGRP 0020
#define MISC_DYNAMIC_MINOR 255
#define VM_IO 0x4000UL
#define VM_PFNMAP 0x400UL
#define VM_DONTEXPAND 0x40000UL
#define VM_DONTDUMP 0x4000000UL

#define READ_ONCE(x) (*(volatile __typeof__(x) *)&(x))
#define WRITE_ONCE(x, v) (*(volatile __typeof__(x) *)&(x) = (v))
#define smp_mb() __sync_synchronize()
#define likely(x) __builtin_expect(!!(x), 1)
#define unlikely(x) __builtin_expect(!!(x), 0)
#define ARRAY_SIZE(a) (sizeof(a) / sizeof((a)[0]))
#define container_of(ptr, type, member) ((type *)((char *)(ptr) - offsetof(type, member)))
#define min_t(t, a, b) ((t)(a) < (t)(b) ? (t)(a) : (t)(b))
#define ALIGN_DOWN(x, a) ((x) & ~((__typeof__(x))(a) - 1))
#define ALIGN(x, a) (((x) + ((__typeof__(x))(a) - 1)) & ~((__typeof__(x))(a) - 1))

/* ---- logging (levels: 0 err, 1 warn, 2 info, 3 debug) */
void sim_log(int level, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
#define pr_err(fmt, ...) sim_log(0, fmt, ##__VA_ARGS__)
#define pr_warn(fmt, ...) sim_log(1, fmt, ##__VA_ARGS__)
#define pr_info(fmt, ...) sim_log(2, fmt, ##__VA_ARGS__)
#define pr_debug(fmt, ...) sim_log(3, fmt, ##__VA_ARGS__)

/* ---- memory */
void *sim_kzalloc(size_t n);
void sim_kfree(const void *p);
#define kzalloc(n, f) sim_kzalloc(n)
#define kmalloc(n, f) sim_kzalloc(n)
#define kcalloc(c, n, f) sim_kzalloc((size_t)(c) * (n))
#define kfree(p) sim_kfree(p)

/* ---- locking */
struct mutex { pthread_mutex_t m; };
static inline void mutex_init(struct mutex *l) {
	pthread_mutexattr_t a;
	pthread_mutexattr_init(&a);
	pthread_mutexattr_settype(&a, PTHREAD_MUTEX_ERRORCHECK); /* a recursive lock is a bug: trap it */
	pthread_mutex_init(&l->m, &a);
}
int sim_mutex_lock(struct mutex *l);
static inline void mutex_lock(struct mutex *l) { sim_mutex_lock(l); }
static inline void mutex_unlock(struct mutex *l) { pthread_mutex_unlock(&l->m); }
static inline void mutex_destroy(struct mutex *l) { pthread_mutex_destroy(&l->m); }
typedef struct { int counter; } atomic_t;
static inline void atomic_set(atomic_t *a, int v) { __atomic_store_n(&a->counter, v, __ATOMIC_SEQ_CST); }
static inline int atomic_read(const atomic_t *a) { return __atomic_load_n(&a->counter, __ATOMIC_SEQ_CST); }
static inline int atomic_inc_return(atomic_t *a) { return __atomic_add_fetch(&a->counter, 1, __ATOMIC_SEQ_CST); }
static inline int atomic_dec_return(atomic_t *a) { return __atomic_sub_fetch(&a->counter, 1, __ATOMIC_SEQ_CST); }
static inline void atomic_inc(atomic_t *a) { __atomic_add_fetch(&a->counter, 1, __ATOMIC_SEQ_CST); }
static inline void atomic_dec(atomic_t *a) { __atomic_sub_fetch(&a->counter, 1, __ATOMIC_SEQ_CST); }
static inline int atomic_dec_and_test(atomic_t *a) { return __atomic_sub_fetch(&a->counter, 1, __ATOMIC_SEQ_CST) == 0; }
typedef struct { long counter; } atomic64_t;
static inline void atomic64_set(atomic64_t *a, long v) { __atomic_store_n(&a->counter, v, __ATOMIC_SEQ_CST); }
static inline void atomic64_inc(atomic64_t *a) { __atomic_add_fetch(&a->counter, 1, __ATOMIC_SEQ_CST); }
static inline void atomic64_dec(atomic64_t *a) { __atomic_sub_fetch(&a->counter, 1, __ATOMIC_SEQ_CST); }
static inline long atomic64_read(const atomic64_t *a) { return __atomic_load_n(&a->counter, __ATOMIC_SEQ_CST); }

/* ---- reference counting (linux/kref.h) */
struct kref { atomic_t refcount; };
static inline void kref_init(struct kref *k) { atomic_set(&k->refcount, 1); }
static inline void kref_get(struct kref *k) { atomic_inc(&k->refcount); }
static inline int kref_put(struct kref *k, void (*release)(struct kref *))
{
	if (atomic_dec_and_test(&k->refcount)) { release(k); return 1; }
	return 0;
}
static inline unsigned int kref_read(const struct kref *k) { return (unsigned int)atomic_read(&k->refcount); }

/* ---- lists */
struct list_head { struct list_head *next, *prev; };
static inline void INIT_LIST_HEAD(struct list_head *h) { h->next = h; h->prev = h; }
static inline void list_add(struct list_head *n, struct list_head *h) {
	n->next = h->next; n->prev = h; h->next->prev = n; h->next = n;
}
static inline void list_add_tail(struct list_head *n, struct list_head *h) {
	n->prev = h->prev; n->next = h; h->prev->next = n; h->prev = n;
}
static inline void list_del(struct list_head *e) {
	e->prev->next = e->next; e->next->prev = e->prev; e->next = e->prev = NULL;
}
static inline int list_empty(const struct list_head *h) { return h->next == h; }
#define list_entry(ptr, type, member) container_of(ptr, type, member)
#define list_for_each_entry(pos, head, member)                                                       \
	for (pos = list_entry((head)->next, __typeof__(*pos), member); &pos->member != (head);          \
	     pos = list_entry(pos->member.next, __typeof__(*pos), member))
#define list_for_each_entry_safe(pos, n, head, member)                                               \
	for (pos = list_entry((head)->next, __typeof__(*pos), member),                                  \
	    n = list_entry(pos->member.next, __typeof__(*pos), member);                                 \
	     &pos->member != (head); pos = n, n = list_entry(n->member.next, __typeof__(*n), member))

/* ---- module plumbing */
struct module { atomic_t refcnt; const char *name; };
extern struct module sim_this_module;
#define THIS_MODULE (&sim_this_module)
static inline void __module_get(struct module *m) { atomic_inc(&m->refcnt); }
static inline void module_put(struct module *m) { atomic_dec(&m->refcnt); }
#define SIM_CAT2(a, b) a##b
#define SIM_CAT(a, b) SIM_CAT2(a, b)
#define module_init(fn) int SIM_CAT(sim_init_, KBUILD_MODNAME)(void) { return fn(); }
#define module_exit(fn) void SIM_CAT(sim_exit_, KBUILD_MODNAME)(void) { fn(); }
#define MODULE_AUTHOR(x)
#define MODULE_LICENSE(x)
#define MODULE_DESCRIPTION(x)
#define MODULE_VERSION(x)
#define MODULE_SOFTDEP(x)
/* A module parameter becomes a setter the tests can call: sim_param_<module>_<name>(value). */
typedef unsigned int uint;
typedef unsigned long ulong;
#define module_param(name, type, perm) \
	__attribute__((visibility("default"))) void SIM_CAT(SIM_CAT(SIM_CAT(sim_param_, KBUILD_MODNAME), _), name)(long v) { name = (type)v; }
#define MODULE_PARM_DESC(name, desc)
#define EXPORT_SYMBOL(x)
size_t strscpy(char *dst, const char *src, size_t n);

/* ---- debugfs + seq_file: one in-memory "file" per debugfs_create_file(); sim_debugfs_read() renders it */
struct dentry { int unused; };
struct seq_file { char *buf; size_t cap, len; void *private; };
void seq_printf(struct seq_file *m, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
struct sim_seq_ops { int (*show)(struct seq_file *, void *); };
#define DEFINE_SHOW_ATTRIBUTE(name) static const struct sim_seq_ops name##_fops = { name##_show }
struct dentry *debugfs_create_dir(const char *name, struct dentry *parent);
struct dentry *debugfs_create_file(const char *name, unsigned short mode, struct dentry *parent, void *data, const struct sim_seq_ops *fops);
void debugfs_remove_recursive(struct dentry *d);
#define IS_ERR_OR_NULL(p) ((p) == NULL)

/* ---- devices */
struct device { int id; const char *name; };
struct pci_dev { struct device dev; unsigned short vendor, device; };
#define to_pci_dev(d) container_of(d, struct pci_dev, dev)
static inline int dev_is_pci(const struct device *d) { return d != NULL; }

/* ---- scatterlists */
struct scatterlist { unsigned long page_link; unsigned int offset, length; dma_addr_t dma_address; unsigned int dma_length; };
struct sg_table { struct scatterlist *sgl; unsigned int nents, orig_nents; };
int sg_alloc_table(struct sg_table *t, unsigned int nents, gfp_t gfp);
void sg_free_table(struct sg_table *t);
static inline struct scatterlist *sg_next(struct scatterlist *sg) { return sg + 1; }
#define for_each_sg(sglist, sg, nr, i) for (i = 0, sg = (sglist); i < (int)(nr); i++, sg = sg_next(sg))
#define sg_dma_address(sg) ((sg)->dma_address)
#define sg_dma_len(sg) ((sg)->dma_length)

/* ---- user copies (fault injection: sim_set_copy_fault(n) fails the n-th copy from now) */
unsigned long sim_copy(void *dst, const void *src, unsigned long n);
#define copy_from_user(d, s, n) sim_copy(d, s, n)
#define copy_to_user(d, s, n) sim_copy(d, s, n)

/* ---- files, misc devices, mmap */
struct inode { int unused; };
struct file { void *private_data; };
struct vm_area_struct { unsigned long vm_start, vm_end, vm_pgoff, vm_flags; pgprot_t vm_page_prot; void *sim_log; };
struct file_operations {
	struct module *owner;
	int (*open)(struct inode *, struct file *);
	int (*release)(struct inode *, struct file *);
	long (*unlocked_ioctl)(struct file *, unsigned int, unsigned long);
	int (*mmap)(struct file *, struct vm_area_struct *);
};
struct miscdevice { int minor; const char *name; const struct file_operations *fops; unsigned short mode; };
int misc_register(struct miscdevice *m);
void misc_deregister(struct miscdevice *m);
int remap_pfn_range(struct vm_area_struct *vma, unsigned long addr, unsigned long pfn, unsigned long size, pgprot_t prot);
#define io_remap_pfn_range remap_pfn_range
static inline pgprot_t pgprot_writecombine(pgprot_t p) { return p | 1; }
static inline void vm_flags_set(struct vm_area_struct *v, unsigned long f) { v->vm_flags |= f; }

#endif
