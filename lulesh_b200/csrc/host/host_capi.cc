// C exports of the host Domain (include/lulesh_host.h) for tests, bench.py and
// foreign-language hosts.
#include <new>
#include <stdexcept>

#include "../../../include/lulesh_host.h"
#include "domain.h"

struct lulesh_host_domain { Domain dom; };

extern "C" lulesh_host_domain *lulesh_host_domain_new(int numRanks, int rank, int px, int py,
                                                      int pz, int sx, int sy, int sz, int numReg,
                                                      int balance, int cost)
{
   try {
      return new lulesh_host_domain{Domain(numRanks, rank, px, py, pz, sx, sy, sz, numReg, balance, cost)};
   } catch (const std::exception &) {
      return nullptr;
   }
}

extern "C" void lulesh_host_domain_free(lulesh_host_domain *d) { delete d; }

extern "C" void lulesh_host_domain_view(lulesh_host_domain *d, lulesh_b200_host_view *out)
{
   *out = d->dom.view();
}

extern "C" double *lulesh_host_domain_field(lulesh_host_domain *d, int field, size_t *count)
{
   std::vector<Real_t> *v = d->dom.realField(field);
   if (count) *count = v ? v->size() : 0;
   return v ? v->data() : nullptr;
}

extern "C" const int32_t *lulesh_host_domain_ints(lulesh_host_domain *d, const char *name, size_t *count)
{
   const std::vector<Index_t> *v = d->dom.intField(name);
   if (count) *count = v ? v->size() : 0;
   return (v && !v->empty()) ? v->data() : nullptr;
}

extern "C" lulesh_b200_scalars *lulesh_host_domain_scalars(lulesh_host_domain *d)
{
   return &d->dom.scalars();
}

extern "C" int lulesh_host_domain_write_vtk(lulesh_host_domain *d, int rank, const char *path)
{
   if (!d || !path) return -1;
   return DumpDomainToVTK(d->dom, rank, path);
}
