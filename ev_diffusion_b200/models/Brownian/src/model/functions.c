/*
 * Brownian random-walk diffusion -- agent scripts (BASELINE config C2).
 *
 * Authored for this repository in the style of the FLAME GPU 1.5 examples
 * (closest relatives: examples/SocialParticleSimulation and
 * examples/CirclesPartitioning_float); it uses only the stock model API:
 * add_<msg>_message, get_first/next_<msg>_message, rnd<CONTINUOUS>.
 *
 * Layers:  output_location -> count_neighbours -> move
 */
#ifndef _FUNCTIONS_H_
#define _FUNCTIONS_H_

#include "header.h"

/* interaction radius for the neighbour count: one partition cell */
#define INTERACTION_RADIUS 1.0f

__FLAME_GPU_FUNC__ int output_location(xmachine_memory_Particle* agent, xmachine_message_location_list* location_messages)
{
	add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z);
	return 0;
}

__FLAME_GPU_FUNC__ int count_neighbours(xmachine_memory_Particle* agent, xmachine_message_location_list* location_messages, xmachine_message_location_PBM* partition_matrix)
{
	const float x1 = agent->x;
	const float y1 = agent->y;
	const float z1 = agent->z;
	int count = 0;

	xmachine_message_location* msg = get_first_location_message(location_messages, partition_matrix, x1, y1, z1);
	while (msg)
	{
		if (msg->id != agent->id)
		{
			const float dx = x1 - msg->x;
			const float dy = y1 - msg->y;
			const float dz = z1 - msg->z;
			const float d2 = dx*dx + dy*dy + dz*dz;
			if (d2 < INTERACTION_RADIUS*INTERACTION_RADIUS)
				count++;
		}
		msg = get_next_location_message(msg, location_messages, partition_matrix);
	}
	agent->nbr = count;
	return 0;
}

/* reflect p back into [0, hi] after a step shorter than the box */
__FLAME_GPU_FUNC__ float reflect(float p, float hi)
{
	if (p < 0.0f) p = -p;
	if (p > hi) p = hi - (p - hi);
	return p;
}

__FLAME_GPU_FUNC__ int move(xmachine_memory_Particle* agent, RNG_rand48* rand48)
{
	const float ux = rnd<CONTINUOUS>(rand48);
	const float uy = rnd<CONTINUOUS>(rand48);
	const float uz = rnd<CONTINUOUS>(rand48);

	float x = agent->x + SIGMA*(2.0f*ux - 1.0f);
	float y = agent->y + SIGMA*(2.0f*uy - 1.0f);
	float z = agent->z + SIGMA*(2.0f*uz - 1.0f);

	agent->x = reflect(x, DOMAIN_X);
	agent->y = reflect(y, DOMAIN_Y);
	agent->z = reflect(z, DOMAIN_Z);
	return 0;
}

#endif
