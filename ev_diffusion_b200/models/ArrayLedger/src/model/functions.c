/*
 * ArrayLedger -- agent scripts.  Authored for this repository; stock model API only
 * (get/set_<A>_agent_array_value, rnd<CONTINUOUS>), in the style of examples/StableMarriage.
 *
 * Layers:  record (death) -> rest (active -> resting) -> wake (resting -> active)
 */
#ifndef _FUNCTIONS_H_
#define _FUNCTIONS_H_

#include "header.h"

/* a random step; the step number and position go into the ring buffers; a few walkers die */
__FLAME_GPU_FUNC__ int record(xmachine_memory_Walker* agent, RNG_rand48* rand48)
{
	const float u = rnd<CONTINUOUS>(rand48);
	agent->x += u - 0.5f;
	set_Walker_agent_array_value<int>(agent->hist, agent->steps % 8, agent->steps * 7 + agent->id);
	set_Walker_agent_array_value<float>(agent->trail, agent->steps % 4, agent->x);
	agent->steps++;
	const float d = rnd<CONTINUOUS>(rand48);
	return d < DEATH_P;
}

/* fold the ring buffer (reads every element of the array after the compaction moved it) */
__FLAME_GPU_FUNC__ int rest(xmachine_memory_Walker* agent)
{
	int s = 0;
	for (unsigned int i = 0; i < xmachine_memory_Walker_hist_LENGTH; i++)
		s += get_Walker_agent_array_value<int>(agent->hist, i) * (int)(i + 1);
	agent->sum = s;
	return 0;
}

__FLAME_GPU_FUNC__ int wake(xmachine_memory_Walker* agent)
{
	const float t = get_Walker_agent_array_value<float>(agent->trail, 3);
	set_Walker_agent_array_value<float>(agent->trail, 3, t + 1.0f);
	return 0;
}

#endif
