/*
 * OptionalSignal -- agent scripts for the optional_message fixture.
 *
 * Authored for this repository with the stock model API only.  A Sensor draws a
 * level; with probability EMIT_P it emits a spatially partitioned `signal`
 * (optional_message output: agents that do not call add_signal_message output
 * nothing).  Every Sensor counts the signals within one cell radius and keeps
 * the loudest emitter's id, then the Sensors that heard a level >= LOUD announce
 * themselves on a non-partitioned optional `beacon`, which everybody tallies.
 * Integer state only, so the parity with the host oracle is exact.
 *
 * Layers:  emit -> listen -> announce -> tally
 */
#ifndef _FUNCTIONS_H_
#define _FUNCTIONS_H_

#include "header.h"

__FLAME_GPU_FUNC__ int emit(xmachine_memory_Sensor* agent, xmachine_message_signal_list* signal_messages, RNG_rand48* rand48)
{
	const float u = rnd<CONTINUOUS>(rand48);
	agent->level = (int)(rnd<CONTINUOUS>(rand48) * 10.0f);
	if (u < EMIT_P)
		add_signal_message(signal_messages, agent->id, agent->x, agent->y, agent->z, agent->level);
	return 0;
}

__FLAME_GPU_FUNC__ int listen(xmachine_memory_Sensor* agent, xmachine_message_signal_list* signal_messages, xmachine_message_signal_PBM* partition_matrix)
{
	int heard = 0;
	int best_level = -1;
	int best_id = -1;
	xmachine_message_signal* msg = get_first_signal_message(signal_messages, partition_matrix, agent->x, agent->y, agent->z);
	while (msg)
	{
		if (msg->id != agent->id)
		{
			const float dx = agent->x - msg->x;
			const float dy = agent->y - msg->y;
			const float dz = agent->z - msg->z;
			if (dx*dx + dy*dy + dz*dz < 1.0f)
			{
				heard++;
				if (msg->level > best_level)   /* ties: the first in message order wins */
				{
					best_level = msg->level;
					best_id = msg->id;
				}
			}
		}
		msg = get_next_signal_message(msg, signal_messages, partition_matrix);
	}
	agent->heard = heard;
	agent->loudest = (best_level >= LOUD) ? best_id : -1;
	return 0;
}

__FLAME_GPU_FUNC__ int announce(xmachine_memory_Sensor* agent, xmachine_message_beacon_list* beacon_messages)
{
	if (agent->loudest >= 0)
		add_beacon_message(beacon_messages, agent->id, agent->heard);
	return 0;
}

__FLAME_GPU_FUNC__ int tally(xmachine_memory_Sensor* agent, xmachine_message_beacon_list* beacon_messages)
{
	int n = 0;
	int last = -1;
	xmachine_message_beacon* msg = get_first_beacon_message(beacon_messages);
	while (msg)
	{
		n += msg->level;
		last = msg->id;            /* order-dependent on purpose: the list must be in agent order */
		msg = get_next_beacon_message(msg, beacon_messages);
	}
	agent->beacons = n * 1024 + (last & 1023);
	return 0;
}

#endif
