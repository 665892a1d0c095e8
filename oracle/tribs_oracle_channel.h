/* TEST INFRASTRUCTURE - see tribs_oracle_channel.c */
#ifndef TRIBS_ORACLE_CHANNEL_H
#define TRIBS_ORACLE_CHANNEL_H

/* The channel network as tKinemat holds it: reaches in the order of NodesLstH / NodesLstO / NNodes
   (tKinemat.cpp:826-830), each a chain of nodes from its head to its outlet (a confluence = the head of the
   next reach, or the basin outlet, cell index n_cells). Per entry: the node's ChannelWidth, the length and slope
   of its flow edge (unused for the last entry of a reach). */
typedef struct {
  int n_cells, n_reach;
  const int *off;        /* n_reach + 1 */
  const int *node;       /* off[n_reach] entries */
  const double *width, *length, *slope;
  double roughness;      /* CHANNELROUGHNESS (uniform, tKinemat.cpp:1221) */
  double uniform_width;  /* > 0: CHANNELWIDTH overrides the per-node widths (tKinemat.cpp:1227-1228) */
} orc_channel;

typedef struct {
  double *hlevel, *qstrm, *velocity;   /* n_cells + 1 (slot n_cells = the basin outlet) */
  double *outlet_hlev;                 /* n_reach: tKinemat::OutletHlev */
} orc_channel_state;

int orc_channel_step(const orc_channel *ch, orc_channel_state *s, const double *qeff, double dt, double *qout, int *iters);

#endif
